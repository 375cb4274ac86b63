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
world = int(os.environ.get('WORLD_SIZE', 1))
rank = int(os.environ.get('RANK', 0))
if world > 1:      # torchrun: data-parallel step with the in-step bucketed all-reduce
    import torch.distributed as dist
    torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', 0)))
    dist.init_process_group('nccl')
dev = torch.device('cuda', torch.cuda.current_device())
torch.manual_seed(0)
head = MHEntHead(mano_data=synthetic_mano(0)).to(dev)
head.q_z_giv_i.precision = os.environ.get('PRECISION', 'bf16x3')
eng = TrainStep(head, B, S, dev, want_verts=True, use_graph=graph, allreduce=world > 1)
eng.load(**{k: v.to(dev) for k, v in synthetic_batch(B, S, seed=1).items()})
for _ in range(5):
    eng.run()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3):
        eng.run()
    torch.cuda.synchronize()
if rank != 0:
    sys.exit(0)
prof.export_chrome_trace('gpurun_out/trace_step.json')
tr = json.load(open('gpurun_out/trace_step.json'))
evs = [e for e in tr['traceEvents'] if e.get('cat') in ('kernel', 'gpu_memset', 'gpu_memcpy')]
evs.sort(key=lambda e: e['ts'])
os.makedirs('gpurun_out', exist_ok=True)
with open('gpurun_out/trace_step.csv', 'w') as fh:
    fh.write('start_us,dur_us,stream,name\n')
    for e in evs:
        fh.write(f"{e['ts']:.3f},{e['dur']:.3f},{e['args'].get('stream', -1)},\"{e['name'][:110]}\"\n")
os.remove('gpurun_out/trace_step.json')
if evs:
    t0, t1 = evs[0]['ts'], max(e['ts'] + e['dur'] for e in evs)
    print(f'kernels {len(evs)} span {(t1 - t0) / 3:.1f} us/step')
