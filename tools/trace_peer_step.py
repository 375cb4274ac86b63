"""Timeline of the data-parallel step with the peer-memory exchange (events recorded inside the captured graph):
torchrun --nproc-per-node N --master-addr 127.0.0.1 tools/trace_peer_step.py   (MHE_FUSED_BWD_CHUNKS etc. apply)"""
import os
import sys

os.environ['MHE_ENGINE_TRACE'] = '1'
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from mhentropy_b200 import MHEntHead
from mhentropy_b200.engine import TrainStep
from mhentropy_b200.mano_assets import synthetic_mano
from mhentropy_b200.synthetic import synthetic_batch

rank, local, world = int(os.environ['RANK']), int(os.environ.get('LOCAL_RANK', 0)), int(os.environ['WORLD_SIZE'])
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
dist.init_process_group('nccl', device_id=dev)
torch.manual_seed(0)
head = MHEntHead(mano_data=synthetic_mano(0)).to(dev)
head.q_z_giv_i.precision = 'bf16x3'
B, S = 64, 10
eng = TrainStep(head, B, S, dev, want_verts=True, use_graph=True, exchange=os.environ.get('MHE_BENCH_EXCHANGE', 'peer'), planes='optimizer')
eng.load(**{k: v.to(dev) for k, v in synthetic_batch(B, S, seed=1000 + rank).items()})
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
rows = {}
for it in range(8):
    flush.zero_()
    torch.cuda.synchronize()
    dist.barrier()
    eng.run()
    if it >= 3:
        for label, ms in eng.trace_ms():
            rows.setdefault(label, []).append(ms * 1e3)
for r in range(world):
    dist.barrier()
    if r == rank and rank in (0, world - 1):
        print(f'--- rank {rank}: us since step start (median of 5 replays)')
        for label, v in rows.items():
            v.sort()
            print(f'  {label:28s} {v[len(v) // 2]:8.1f}')
        sys.stdout.flush()
dist.barrier()
dist.destroy_process_group()
