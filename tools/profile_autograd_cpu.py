"""Host-side profile (cProfile) of the autograd drop-in step: MHEntHead.get_loss + backward on the bench shape."""
import cProfile
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from mhentropy_b200 import MHEntHead
from mhentropy_b200.mano_assets import synthetic_mano
from mhentropy_b200.synthetic import synthetic_batch

dev = torch.device('cuda')
B, S = 64, 10
head = MHEntHead(mano_data=synthetic_mano(0)).to(dev)
for p in head.parameters():
    p.requires_grad_(True)
plist = list(head.parameters())
host = {k: v.pin_memory() for k, v in synthetic_batch(B, S, seed=1).items()}


def step():
    feat = host['feat'].to(dev, non_blocking=True).requires_grad_(True)
    z_det = host['z_det'].to(dev, non_blocking=True).requires_grad_(True)
    z0 = host['z0'].to(dev, non_blocking=True)
    y = {'crop_uv': host['crop_uv'].to(dev, non_blocking=True), 'vis': host['vis'].to(dev, non_blocking=True)}
    for p in plist:
        p.grad = None
    out = head.get_loss(feat, y, z0=z0, z_det=z_det, N=S, want_verts=True)
    loss = (-out['log_p']).mean()
    loss.backward()
    return loss.item()


for _ in range(10):
    step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(100):
    step()
torch.cuda.synchronize()
print(f'wall {1e3 * (time.perf_counter() - t0) / 100:.3f} ms/step')
# host time alone: no .item() sync inside
def step_nosync():
    feat = host['feat'].to(dev, non_blocking=True).requires_grad_(True)
    z_det = host['z_det'].to(dev, non_blocking=True).requires_grad_(True)
    z0 = host['z0'].to(dev, non_blocking=True)
    y = {'crop_uv': host['crop_uv'].to(dev, non_blocking=True), 'vis': host['vis'].to(dev, non_blocking=True)}
    for p in plist:
        p.grad = None
    out = head.get_loss(feat, y, z0=z0, z_det=z_det, N=S, want_verts=True)
    (-out['log_p']).mean().backward()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(100):
    step_nosync()
t1 = time.perf_counter()
torch.cuda.synchronize()
print(f'host enqueue {1e3 * (t1 - t0) / 100:.3f} ms/step, with drain {1e3 * (time.perf_counter() - t0) / 100:.3f} ms/step')
pr = cProfile.Profile()
pr.enable()
for _ in range(200):
    step_nosync()
pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats('cumulative').print_stats(28)
