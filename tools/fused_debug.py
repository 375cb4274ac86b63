"""Phase timeline of the cluster-fused flow kernel (CTA 0): MHE_FUSED_DEBUG=1 makes it stamp %globaltimer at phase
boundaries.  usage: MHE_FUSED_DEBUG=1 python tools/fused_debug.py [B] [S] [bwd]"""
import ctypes
import os
import sys

os.environ.setdefault('MHE_FUSED_DEBUG', '1')
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from mhentropy_b200 import RealNVP
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
flat, packed = flow.flat_parameters(dev), flow.packed_weights(dev)
feat, z0 = torch.randn(B, 512, device=dev), torch.randn(R, 45, device=dev)
with torch.no_grad():
    cp = flow.cond_projections(feat)
x, logdet = torch.empty(R, 45, device=dev), torch.empty(R, device=dev)
wsb = L.mhe_flow_workspace_bytes(shape, R, 1)
ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
for _ in range(3):
    check(L.mhe_flow_pass_fwd(shape, ptr(flat), ptr(packed), ptr(flow.mask), ptr(cp), ptr(z0), R, B, 0, ptr(x), ptr(logdet), None, ptr(ws), wsb,
                              stream_ptr(dev)), 'fwd')
torch.cuda.synchronize()
n = 12 * 64
buf = (ctypes.c_longlong * n)()
L.mhe_fused_debug_read.argtypes = [ctypes.c_void_p, ctypes.c_int]
check(L.mhe_fused_debug_read(buf, n), 'debug read')
t0 = buf[0]
names = {0: 'W step start', 1: 'W cp0 issued', 2: 'W acc0 ready', 3: 'W a0 stored', 4: 'W a0 signalled', 5: 'W cp1 issued', 6: 'W acc1 ready',
         7: 'W a1 in smem', 8: 'W acc2 ready', 9: 'W partial signalled', 10: 'W partials ready', 11: 'W coupling done',
         16: 'P before a0 wait', 17: 'P a0 ready', 20: 'M W0 landed', 21: 'M xm ready', 22: 'M own a0 ready',
         24: 'M W2 landed', 25: 'M a1 ready'}
names.update({32 + i: f'M   A block {i} landed' for i in range(8)})
names.update({48 + i: f'PA  issue W1 block {i}' for i in range(8)})
names.update({56: 'PA  issue W0', 57: 'PA  issue W2'})
names.update({40 + i: f'M   B block {i} landed' for i in range(8)})
for step in range(12):
    ev = sorted((buf[step * 64 + k], k) for k in names if buf[step * 64 + k])
    print(f'--- step {step}')
    for ts, k in ev:
        print(f'   {(ts - t0) / 1.8e3:9.2f} us  {names[k]}')

for step in range(1, 12):
    dcy = buf[step * 64 + 0] - buf[(step - 1) * 64 + 0]
    dns = buf[step * 64 + 60] - buf[(step - 1) * 64 + 60]
    print(f'step {step}: {dns} ns, {dcy} cycles -> {dcy / max(dns, 1):.3f} GHz')
