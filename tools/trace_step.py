"""Kernel timeline of one fused training step (CUPTI through torch.profiler): per-kernel start/duration/stream,
so the dependent-launch critical path and the gaps between kernels are visible.  Writes gpurun_out/trace_step.csv."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from mhentropy_b200 import MHEntHead
from mhentropy_b200.engine import TrainStep
from mhentropy_b200.mano_assets import synthetic_mano
from mhentropy_b200.synthetic import synthetic_batch

B, S = int(os.environ.get('B', 64)), int(os.environ.get('S', 10))
graph = os.environ.get('GRAPH', '1') == '1'
dev = torch.device('cuda')
torch.manual_seed(0)
head = MHEntHead(mano_data=synthetic_mano(0)).to(dev)
head.q_z_giv_i.precision = os.environ.get('PRECISION', 'bf16x3')
eng = TrainStep(head, B, S, dev, want_verts=True, use_graph=graph)
eng.load(**{k: v.to(dev) for k, v in synthetic_batch(B, S, seed=1).items()})
for _ in range(5):
    eng.run()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3):
        eng.run()
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
rows = sorted(((e.time_range.start, e.time_range.end - e.time_range.start, e.name) for e in evs), key=lambda r: r[0])
os.makedirs('gpurun_out', exist_ok=True)
with open('gpurun_out/trace_step.csv', 'w') as fh:
    fh.write('start_us,dur_us,name\n')
    for st, du, nm in rows:
        fh.write(f'{st:.3f},{du:.3f},"{nm[:120]}"\n')
if rows:
    t0, t1 = rows[0][0], max(r[0] + r[1] for r in rows)
    busy = sum(r[1] for r in rows)
    print(f'kernels {len(rows)} span {(t1 - t0) / 3:.1f} us/step, sum of kernel durations {busy / 3:.1f} us/step')
