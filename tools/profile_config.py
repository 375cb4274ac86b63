"""One step of a BASELINE configuration, for launch lists / ncu captures (no CUDA graph: every kernel is its own launch).
usage: CONFIG=5shard|3|4|2 [STEPS=1] python tools/profile_config.py      (under ncu: --metrics gpu__time_duration.sum ...)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from mhentropy_b200 import MHEntHead
from mhentropy_b200.engine import TrainStep
from mhentropy_b200.mano_assets import synthetic_mano
from mhentropy_b200.synthetic import synthetic_batch

cfg = os.environ.get('CONFIG', '5shard')
steps = int(os.environ.get('STEPS', 1))
dev = torch.device('cuda')
torch.manual_seed(0)
head = MHEntHead(mano_data=synthetic_mano(0)).to(dev)
if cfg in ('5shard', '2'):
    B, S = (512, 64) if cfg == '5shard' else (64, 10)
    eng = TrainStep(head, B, S, dev, want_verts=True, use_graph=False, planes='optimizer')
    eng.load(**{k: v.to(dev) for k, v in synthetic_batch(B, S, seed=1).items()})
    fn = eng.run
elif cfg == '3':
    B, S = 256, 100
    feat, z0 = torch.randn(B, 512, device=dev), torch.randn(B * S, 45, device=dev) * 0.8
    z_det = torch.cat([0.5 * torch.randn(B, 3), 0.02 * torch.randn(B, 10), torch.randn(B, 1) * 0.1 - 1.2, 0.1 * torch.randn(B, 2)], 1).to(dev)
    fn = lambda: head.sample(feat, N=S, temp=0.8, mods={'uv', 'xyz', 'verts'}, z0=z0, z_det=z_det)  # noqa: E731
else:
    R = 16384
    x, featr = 0.5 * torch.randn(R, 45, device=dev), torch.randn(R, 512, device=dev)

    def fn():
        with torch.no_grad():
            return head.q_z_giv_i.log_prob(x, logvar=featr)
fn()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(steps):
    fn()
b.record()
torch.cuda.synchronize()
print(f'config {cfg}: {a.elapsed_time(b) / steps:.3f} ms/step (stream launches, warm)')
