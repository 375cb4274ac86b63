"""MANO mesh forward alone (mhe_mano_fwd with vertices) at a given row count: ms, rows/s, output GB/s.  ROWS=25600 python tools/bench_mesh.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from mhentropy_b200 import MHEntHead
from mhentropy_b200._lib import check, lib, ptr, stream_ptr
from mhentropy_b200.mano_assets import synthetic_mano

dev = torch.device('cuda')
R = int(os.environ.get('ROWS', 25600))
torch.manual_seed(0)
head = MHEntHead(mano_data=synthetic_mano(0)).to(dev)
consts = head.mano_dec.mano_layer._consts(dev)
zz = torch.cat([0.5 * torch.randn(R, 3), 1.0 * torch.randn(R, 45), 0.02 * torch.randn(R, 10), torch.zeros(R, 3)], 1).to(dev).contiguous()
verts, jtr = torch.empty(R, 778, 3, device=dev), torch.empty(R, 21, 3, device=dev)
wsb = lib().mhe_mano_workspace_bytes(R, 0)
ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def mesh():
    check(lib().mhe_mano_fwd(consts, zz.data_ptr(), 61, zz.data_ptr() + 48 * 4, 61, R, 1, ptr(verts), ptr(jtr), None, ptr(ws), wsb, stream_ptr(dev)), 'mano_fwd')


for _ in range(3):
    mesh()
torch.cuda.synchronize()
tot = 0.0
n = 10
for _ in range(n):
    flush.zero_()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    mesh()
    b.record()
    torch.cuda.synchronize()
    tot += a.elapsed_time(b)
ms = tot / n
print(f'mesh forward {R} rows: {ms * 1e3:.1f} us, {R / ms / 1e3:.2f} M rows/s, vertices written at {R * 9336 / ms / 1e6:.0f} GB/s '
      f'(MHE_MANO_TC_SKIN={os.environ.get("MHE_MANO_TC_SKIN")}, MHE_MANO_FUSED={os.environ.get("MHE_MANO_FUSED")})')
