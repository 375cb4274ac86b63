"""Time one flow pass (forward, no saving / with saving, and backward) through the C ABI at a given B x S.
MHE_FUSED_MAX_ROWS=0 selects the per-GEMM tensor-core path for comparison.
usage: python tools/bench_flow_pass.py [B] [S]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from mhentropy_b200 import RealNVP, _lib
from mhentropy_b200._lib import check, lib, ptr, stream_ptr

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
S = int(sys.argv[2]) if len(sys.argv) > 2 else 10
R = B * S
dev = torch.device('cuda')
torch.manual_seed(0)
flow = RealNVP(dim=45, tsfm_on=512, h_dims=[512, 512], num_steps=6).to(dev)
flow.precision = 'bf16x3'
L = lib()
shape = flow._shape
flat = flow.flat_parameters(dev)
packed = flow.packed_weights(dev)
feat = torch.randn(B, 512, device=dev)
z0 = torch.randn(R, 45, device=dev)
with torch.no_grad():
    cp = flow.cond_projections(feat)
x = torch.empty(R, 45, device=dev)
logdet = torch.empty(R, device=dev)
wsb = L.mhe_flow_workspace_bytes(shape, R, 1)
ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
saved = torch.empty(L.mhe_flow_saved_bytes(shape, R, 1), dtype=torch.uint8, device=dev)
dflat = torch.zeros_like(flat)
dcp = torch.zeros_like(cp)
dx = torch.randn(R, 45, device=dev)
dld = torch.randn(R, device=dev)
dz0 = torch.empty(R, 45, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
s = stream_ptr(dev)


def fwd(save):
    check(L.mhe_flow_pass_fwd(shape, ptr(flat), ptr(packed), ptr(flow.mask), ptr(cp), ptr(z0), R, B, 0, ptr(x), ptr(logdet),
                              ptr(saved) if save else None, ptr(ws), wsb, s), 'fwd')


def bwd():
    check(L.mhe_flow_pass_bwd(shape, ptr(flat), ptr(packed), ptr(flow.mask), ptr(cp), ptr(saved), R, B, 0, ptr(dx), ptr(dld), -1.0,
                              ptr(dz0), ptr(dflat), ptr(dcp), ptr(ws), wsb, s), 'bwd')


def timeit(fn, n=20, cold=True):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(n):
        if cold:
            flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        tot += a.elapsed_time(b)
    return tot / n * 1e3


print(f'B={B} S={S} R={R} fused_max_rows={os.environ.get("MHE_FUSED_MAX_ROWS", "default")}')
print(f'  fwd (no save)  cold {timeit(lambda: fwd(False)):8.1f} us   warm {timeit(lambda: fwd(False), cold=False):8.1f} us')
print(f'  fwd (save)     cold {timeit(lambda: fwd(True)):8.1f} us   warm {timeit(lambda: fwd(True), cold=False):8.1f} us')
fwd(True)
print(f'  bwd            cold {timeit(bwd):8.1f} us   warm {timeit(bwd, cold=False):8.1f} us')
