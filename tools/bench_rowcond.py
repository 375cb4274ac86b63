"""NLL scoring with one feature row per pose (BASELINE configs[3]): the conditioning contracted inside the coupling GEMMs
(mhe_flow_pass_fwd_rowcond) against the pass on materialised projections (MHE_FLOW_ROWCOND=0).  L2 flushed before every call."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from mhentropy_b200 import RealNVP
from oracle import flow_oracle as fo

DEV = 'cuda'
flow = RealNVP(dim=45, tsfm_on=512, kemb=False, jointN=21, h_dims=[512, 512], num_steps=6)
flow.precision = 'bf16x3'
flow.load_state_dict(fo.init_state_dict(seed=0), strict=True)
flow = flow.to(DEV)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
g = torch.Generator().manual_seed(3)
for R in [int(r) for r in os.environ.get('ROWS', '8192,16384,65536').split(',')]:
    x, feat = (0.5 * torch.randn(R, 45, generator=g)).to(DEV), torch.randn(R, 512, generator=g).to(DEV)
    res = {}
    for mode in ('0', '1', '0', '1'):
        os.environ['MHE_FLOW_ROWCOND'] = mode
        with torch.no_grad():
            for _ in range(2):
                lp = flow.log_prob(x, logvar=feat)
            ts = []
            for _ in range(5):
                flush.zero_()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                lp = flow.log_prob(x, logvar=feat)
                b.record()
                torch.cuda.synchronize()
                ts.append(a.elapsed_time(b))
        ts.sort()
        res.setdefault(mode, []).append(ts[len(ts) // 2])
        res['lp' + mode] = lp
    err = float((res['lp0'] - res['lp1']).abs().max() / res['lp0'].abs().max())
    print(f'R={R}: materialised {min(res["0"]):.3f} ms, in-GEMM conditioning {min(res["1"]):.3f} ms, log_prob rel difference {err:.2e}', flush=True)
