"""Gradient all-reduce candidates on N GPUs (torchrun): NCCL vs the symmetric-memory (NVLink P2P / NVLS multimem) kernels of
torch.distributed._symmetric_memory, on the 80 MB flat gradient.  usage: torchrun --nproc-per-node N tools/bench_allreduce.py"""
import os
import sys

import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem

rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ.get('LOCAL_RANK', 0))
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
dist.init_process_group('nccl')
n = 20_030_592                      # floats of the flat gradient (padded; divisible by 8 ranks x 4 floats)
group = dist.group.WORLD
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    ts = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    t = torch.tensor([ts[len(ts) // 2]], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


x = torch.randn(n, device=dev)
ref = x.clone()
dist.all_reduce(ref)
res = {'nccl': timeit(lambda: dist.all_reduce(x))}
try:
    t = symm_mem.empty(n, dtype=torch.float32, device=dev)
    hdl = symm_mem.rendezvous(t, group.group_name)
    for name in ('two_shot_all_reduce_', 'multimem_all_reduce_'):
        try:
            op = getattr(torch.ops.symm_mem, name)
            t.copy_(ref / world if False else torch.randn(n, device=dev, generator=torch.Generator(device=dev).manual_seed(rank)))
            want = t.clone()
            dist.all_reduce(want)
            op(t, 'sum', group.group_name)
            torch.cuda.synchronize()
            err = float((t - want).abs().max() / want.abs().max())
            res[name] = timeit(lambda: op(t, 'sum', group.group_name))
            res[name + '_relerr'] = err
        except Exception as e:  # noqa: BLE001
            res[name] = 'failed: ' + str(e)[:120]
    res['multicast'] = bool(hdl.has_multicast_support) if hasattr(hdl, 'has_multicast_support') else None
    # copy-engine exchange: push shard p of my gradient into peer p's landing buffer, barrier, sum, push the reduced shard back
    try:
        sh = n // world
        land = symm_mem.empty(world * sh, dtype=torch.float32, device=dev)
        hl = symm_mem.rendezvous(land, group.group_name)
        peers = [(rank + i) % world for i in range(1, world)]
        land_of = {p: hl.get_buffer(p, (world, sh), torch.float32) for p in peers}
        grad_of = {p: hdl.get_buffer(p, (world, sh), torch.float32) for p in peers}
        mine = t.view(world, sh)
        streams = [torch.cuda.Stream(dev) for _ in peers]

        def ce_all_reduce(n_streams=len(peers)):
            main = torch.cuda.current_stream()
            hdl.barrier(channel=0)                      # every rank's gradient is complete, last round's landing buffers are consumed
            for i, p in enumerate(peers):
                st = streams[i % n_streams]
                st.wait_stream(main)
                with torch.cuda.stream(st):
                    land_of[p][rank].copy_(mine[p], non_blocking=True)
            for st in streams[:n_streams]:
                main.wait_stream(st)
            hdl.barrier(channel=1)                      # every push has landed
            lv = land.view(world, sh)
            acc = mine[rank]
            for p in peers:
                acc.add_(lv[p])
            for i, p in enumerate(peers):
                st = streams[i % n_streams]
                st.wait_stream(main)
                with torch.cuda.stream(st):
                    grad_of[p][rank].copy_(acc, non_blocking=True)
            for st in streams[:n_streams]:
                main.wait_stream(st)
            hdl.barrier(channel=2)

        t.copy_(torch.randn(n, device=dev, generator=torch.Generator(device=dev).manual_seed(rank)))
        want = t.clone()
        dist.all_reduce(want)
        torch.cuda.synchronize()
        ce_all_reduce()
        torch.cuda.synchronize()
        res['ce_relerr'] = float((t - want).abs().max() / want.abs().max())
        res['ce_all_reduce'] = timeit(ce_all_reduce)
        if world > 2:
            res['ce_all_reduce_1stream'] = timeit(lambda: ce_all_reduce(1))
        g = torch.cuda.CUDAGraph()
        try:
            with torch.cuda.graph(g):
                ce_all_reduce()
            res['ce_all_reduce_graph'] = timeit(g.replay)
        except Exception as e:  # noqa: BLE001
            res['ce_all_reduce_graph'] = 'failed: ' + str(e)[:160]
    except Exception as e:  # noqa: BLE001
        res['ce_all_reduce'] = 'failed: ' + str(e)[:300]
except Exception as e:  # noqa: BLE001
    res['symm_mem'] = 'failed: ' + str(e)[:200]
if rank == 0:
    print(world, 'GPUs, 80 MB fp32, median us:', res)
dist.destroy_process_group()
