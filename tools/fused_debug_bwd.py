"""Phase timeline of the cluster-fused BACKWARD kernel (CTA 0, clock64 stamps): one training step through the engine with
MHE_FUSED_DEBUG=1 MHE_FUSED_DEBUG_BWD=1 MHE_FUSED_BWD_CHUNKS=1.  usage: python tools/fused_debug_bwd.py [B] [S]"""
import ctypes
import os
import sys

os.environ['MHE_FUSED_DEBUG'] = '1'
os.environ['MHE_FUSED_DEBUG_BWD'] = '1'
os.environ['MHE_FUSED_BWD_CHUNKS'] = '1'
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from mhentropy_b200 import MHEntHead
from mhentropy_b200._lib import check, lib
from mhentropy_b200.engine import TrainStep
from mhentropy_b200.mano_assets import synthetic_mano
from mhentropy_b200.synthetic import synthetic_batch

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
S = int(sys.argv[2]) if len(sys.argv) > 2 else 10
dev = torch.device('cuda')
torch.manual_seed(0)
head = MHEntHead(mano_data=synthetic_mano(0)).to(dev)
head.q_z_giv_i.precision = 'bf16x3'
eng = TrainStep(head, B, S, dev, want_verts=True, use_graph=False)
eng.load(**{k: v.to(dev) for k, v in synthetic_batch(B, S, seed=1).items()})
for _ in range(3):
    eng.run()
torch.cuda.synchronize()
L = lib()
n = 12 * 64
buf = (ctypes.c_longlong * n)()
L.mhe_fused_debug_read.argtypes = [ctypes.c_void_p, ctypes.c_int]
check(L.mhe_fused_debug_read(buf, n), 'debug read')
names = {0: 'step start', 1: 'dpre tile ready', 2: 'acc0 (bG2) ready', 3: 'dh1 slice stored', 4: 'dh1 signalled', 6: 'acc1 (bG1) ready',
         7: 'dh0 stored + copied', 8: 'acc2 (bG0) ready', 9: 'partials signalled', 10: 'partials ready', 11: 'coupling of next layer done'}
t0 = buf[0]
for it in range(12):
    ev = sorted((buf[it * 64 + k], k) for k in names if buf[it * 64 + k])
    print(f'--- iteration {it}')
    prev = None
    for ts, k in ev:
        print(f'   {(ts - t0) / 1.9e3:9.2f} us  (+{0 if prev is None else (ts - prev) / 1.9e3:5.2f})  {names[k]}')
        prev = ts
