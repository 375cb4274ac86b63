"""Time the conditioning GEMM (mhe_flow_cond_fwd) alone: usage  python tools/bench_cond.py [B] [num_steps]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from mhentropy_b200 import RealNVP

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
dev = torch.device('cuda')
torch.manual_seed(0)
flow = RealNVP(dim=45, tsfm_on=512, h_dims=[512, 512], num_steps=steps).to(dev)
flow.precision = 'bf16x3'
feat = torch.randn(B, 512, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
with torch.no_grad():
    for _ in range(3):
        cp = flow.cond_projections(feat)
    ts = []
    for _ in range(20):
        if os.environ.get('NOFLUSH', '0') != '1':
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        cp = flow.cond_projections(feat)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
ts.sort()
wbytes = steps * 2 * 4 * 512 * 512 * 4
print(f'B={B} layers={steps * 2} direct={os.environ.get("MHE_COND_DIRECT", "1")}: median {ts[len(ts) // 2]:.1f} us (min {ts[0]:.1f}); '
      f'fp32 weights {wbytes / 1e6:.1f} MB -> {wbytes / ts[len(ts) // 2] / 1e6:.2f} TB/s')
