"""Data-parallel gradient exchange: factored (all-gather of the conditioning factors + local mhe_flow_cond_wgrad, all-reduce of the rest)
against one dense all-reduce of the flat gradient.  usage: torchrun --nproc-per-node N tools/check_factored_exchange.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from mhentropy_b200 import MHEntHead
from mhentropy_b200.engine import TrainStep
from mhentropy_b200.mano_assets import synthetic_mano
from mhentropy_b200.synthetic import synthetic_batch

rank, local = int(os.environ['RANK']), int(os.environ.get('LOCAL_RANK', 0))
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
dist.init_process_group('nccl')
torch.manual_seed(0)
head = MHEntHead(mano_data=synthetic_mano(0)).to(dev)
head.q_z_giv_i.precision = 'bf16x3'
B, S = 64, 10
devb = {k: v.to(dev) for k, v in synthetic_batch(B, S, seed=100 + rank).items()}
out = []
for factored in (False, True):
    eng = TrainStep(head, B, S, dev, want_verts=False, use_graph=True, factored_exchange=factored)
    eng.load(**devb)
    for _ in range(2):
        eng.run()
        eng.exchange_gradients()
    torch.cuda.synchronize()
    out.append((eng.dflat.clone(), eng.loss.clone()))
a, b = out
err = float((a[0] - b[0]).abs().max() / a[0].abs().max())
rel_fro = float((a[0].double() - b[0].double()).norm() / a[0].double().norm())
if rank == 0:
    print(f'world {dist.get_world_size()}: dense vs factored exchange: max-rel {err:.3e}, frobenius-rel {rel_fro:.3e}, loss {float(a[1]):.6f} / {float(b[1]):.6f}')
assert rel_fro < 1e-5 and abs(float(a[1]) - float(b[1])) < 1e-3 * abs(float(a[1]))
dist.destroy_process_group()
