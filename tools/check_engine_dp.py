"""Data-parallel TrainStep == single-process TrainStep on the concatenated batch (global-batch-mean scaling).

Every rank runs the engine on its shard (images [g*B, (g+1)*B) of a global batch of world*B images, all S hypotheses) and exchanges
the gradients (factored, dense, and whatever 'auto' picks); rank 0 also runs ONE engine on the whole batch.  Loss, flat gradient and the
per-image feature gradient must agree: each rank seeds its backward with 1/(B*world) and the exchange sums.
usage: torchrun --nproc-per-node N --master-addr 127.0.0.1 tools/check_engine_dp.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from mhentropy_b200 import MHEntHead
from mhentropy_b200.engine import TrainStep
from mhentropy_b200.mano_assets import synthetic_mano
from mhentropy_b200.parallel import shard_batch
from mhentropy_b200.synthetic import synthetic_batch

rank, local, world = int(os.environ['RANK']), int(os.environ.get('LOCAL_RANK', 0)), int(os.environ['WORLD_SIZE'])
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
dist.init_process_group('nccl', device_id=dev)
torch.manual_seed(0)
head = MHEntHead(mano_data=synthetic_mano(0)).to(dev)
head.q_z_giv_i.precision = 'bf16x3'
B, S = 16, 10
full = synthetic_batch(B * world, S, seed=321)
mine = {k: v.to(dev) for k, v in shard_batch(full, S, rank, world).items()}
fro = lambda a, b: float((a.double() - b.double()).norm() / (b.double().norm() + 1e-300))  # noqa: E731
ok = True
# the peer-memory exchange by itself against NCCL on ragged ranges (unaligned starts, a 1-float range, an untouched gap), eager and captured
from mhentropy_b200.parallel import PeerExchange
n = 1_000_003
px = PeerExchange(n, dev).plan([[(0, 1000), (1003, 50_001)], [(50_001, 50_002), (60_000, n)], [(px_tail := 1_000_032, px_tail + 1)]])
for use_graph in (False, True):
    g = torch.cuda.CUDAGraph() if use_graph else None
    if g is not None:
        with torch.cuda.graph(g):
            px.all_reduce()
    for it in range(3):
        src = torch.randn(px.size, device=dev, generator=torch.Generator(device=dev).manual_seed(100 * it + rank))
        px.buf.copy_(src)
        want = src.clone()
        dist.all_reduce(want)
        want[50_002:60_000] = src[50_002:60_000]        # not exchanged
        want[1000:1003] = src[1000:1003]
        want[n:px_tail] = src[n:px_tail]
        want[px_tail + 1:] = src[px_tail + 1:]
        torch.cuda.synchronize()
        dist.barrier()
        g.replay() if g is not None else px.all_reduce()
        torch.cuda.synchronize()
        err = float((px.buf - want).abs().max())
        if err > 1e-5:
            ok = False
        if rank == 0:
            print(f'PeerExchange vs NCCL (graph={use_graph}, round {it}): max abs error {err:.2e}')
del px, g
results = {}
for mode in ('dense', 'factored', 'auto', 'peer'):
    eng = TrainStep(head, B, S, dev, want_verts=False, use_graph=True, exchange=mode)
    eng.load(**mine)
    for _ in range(4 if mode == 'peer' else 2):
        eng.run()
        eng.exchange_gradients()
    torch.cuda.synchronize()
    results[mode] = (eng.dflat.clone(), eng.loss.clone(), eng.dfeat.clone(), eng.factored_exchange)
    del eng
if rank == 0:
    # the reference point: one process, the whole batch (no process group involved: world = 1 semantics through a sub-engine)
    single = TrainStep(head, B * world, S, dev, want_verts=False, use_graph=False, average_over_ranks=False, exchange='dense')
    single.world = 1
    single.load(**{k: v.to(dev) for k, v in full.items()})
    single.run()
    torch.cuda.synchronize()
    for mode, (dflat, loss, dfeat, factored) in results.items():
        e_g, e_l = fro(dflat, single.dflat), abs(float(loss) - float(single.loss)) / abs(float(single.loss))
        e_f = fro(dfeat, single.dfeat[:B])            # rank 0 owns the first B images; dfeat is per image (not reduced)
        print(f'world {world} exchange={mode} (factored={factored}): flat gradient vs single process {e_g:.2e}, loss {e_l:.2e}, dfeat[rank 0] {e_f:.2e}')
        ok = ok and e_g < 2e-4 and e_l < 1e-5 and e_f < 2e-4
flag = torch.tensor([1 if ok else 0], device=dev)
dist.broadcast(flag, 0)
dist.barrier()
dist.destroy_process_group()
assert int(flag) == 1
if rank == 0:
    print('check_engine_dp: OK')
