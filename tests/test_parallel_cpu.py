"""N > 1 path on CPU (gloo, world_size 2): sharding by image + all-reduce reproduces the single-process step."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mhentropy_b200.mano_assets import synthetic_mano
from mhentropy_b200.parallel import allreduce_step, cond_wgrad_from_factors, gather_cond_factors, image_range, shard_batch
from mhentropy_b200.synthetic import synthetic_batch
from oracle import flow_oracle as fo
from oracle import loss_oracle as lo
from oracle import mano_oracle as mo

CFG = dict(dim=45, cond_dim=32, h_dims=(64, 64), num_steps=2)
B, S = 5, 3     # odd image count: ragged shards


def _step(batch, n_images_local):
    sd = fo.init_state_dict(seed=4, **CFG)
    sdg = {k: v.clone().requires_grad_(k != 'mask') for k, v in sd.items()}
    c = mo.mano_constants(synthetic_mano(0))
    out = lo.reverse_kld(sdg, c, batch['feat'], batch['z_det'], batch['z0'], batch['crop_uv'], batch['vis'], S)
    loss = lo.mhent_loss(out['log_p'])
    loss.backward()
    flat = torch.cat([v.grad.flatten() for k, v in sdg.items() if k != 'mask'])
    return loss.detach().reshape(1).clone(), flat


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    torch.set_num_threads(1)
    full = synthetic_batch(B, S, seed=9, cond_dim=32)
    mine = shard_batch(full, S, rank, world)
    lo_, hi_ = image_range(B, rank, world)
    loss, flat = _step(mine, hi_ - lo_)
    allreduce_step(flat, loss, hi_ - lo_, B)
    if rank == 0:
        q.put((loss.numpy(), flat.numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_image_range_partitions():
    for Bn in (1, 5, 64, 4096):
        for w in (1, 2, 3, 8):
            spans = [image_range(Bn, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == Bn
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


def test_shard_keeps_hypothesis_major_rows():
    full = synthetic_batch(B, S, seed=9, cond_dim=32)
    sh = shard_batch(full, S, 1, 2)
    lo_, hi_ = image_range(B, 1, 2)
    nb = hi_ - lo_
    for n in range(S):
        for j in range(nb):
            assert torch.equal(sh['z0'][n * nb + j], full['z0'][n * B + lo_ + j])
    assert torch.equal(sh['feat'], full['feat'][lo_:hi_])


def test_two_rank_step_matches_single_process():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    loss2, flat2 = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    full = synthetic_batch(B, S, seed=9, cond_dim=32)
    loss1, flat1 = _step(full, B)
    assert abs(float(loss2[0]) - float(loss1)) < 1e-4 * abs(float(loss1))
    err = float((torch.from_numpy(flat2) - flat1).norm() / flat1.norm())
    assert err < 1e-4, err


def _factor_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    torch.set_num_threads(1)
    g = torch.Generator().manual_seed(50 + rank)
    Bl, L4, H, C = 3, 8, 16, 12
    dcp, feat = torch.randn(Bl, L4 * H, generator=g, dtype=torch.float64), torch.randn(Bl, C, generator=g, dtype=torch.float64)
    dense = cond_wgrad_from_factors(dcp, feat, H)             # the local gradient ...
    dist.all_reduce(dense)                                     # ... all-reduced: the dense exchange
    dcp_all, feat_all = gather_cond_factors(dcp, feat)         # the factored exchange
    fact = cond_wgrad_from_factors(dcp_all, feat_all, H)
    if rank == 0:
        q.put((dense.numpy(), fact.numpy(), tuple(dcp_all.shape)))
    dist.barrier()
    dist.destroy_process_group()


def test_factored_cond_exchange_equals_dense_allreduce():
    """The data-parallel exchange of the conditioning weight gradient by its factors (all-gather of dcp / feat, then one local contraction)
    equals the all-reduce of the local dense gradients (gloo, world size 2)."""
    import numpy as np
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_factor_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    dense, fact, shape = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert shape == (6, 8 * 16)
    np.testing.assert_allclose(fact, dense, rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize('world', [2, 3, 8])
def test_peer_exchange_plan_reduces_every_range_once(world):
    """The shard plan of the peer-memory exchange (parallel.PeerExchange), replayed on the host with one array per rank: pushes into the
    landing slots, the owner's sum, pushes back.  Exchanged ranges end up as the sum over the ranks on every rank, everything else is
    untouched, no landing slot is used twice (ragged ranges: unaligned starts, a 1-float range, an empty one)."""
    import numpy as np
    from mhentropy_b200.parallel import plan_peer_buckets, shard_length, shard_range
    n = 5000
    buckets = [[(0, 1000), (1003, 2001)], [(2001, 2002), (2500, 2500)], [(2600, n)]]
    plan, land_floats = plan_peer_buckets(buckets, world, align=32)
    assert shard_length(1, world) == 32 and shard_length(0, world) == 0
    rng = np.random.default_rng(0)
    src = [rng.standard_normal(n) for _ in range(world)]
    buf = [s.copy() for s in src]
    land = [np.full((world, land_floats), np.nan) for _ in range(world)]
    used = np.zeros(land_floats, dtype=int)
    for segs in plan:
        for seg in segs:
            los = [shard_range(seg, p) for p in range(world)]
            assert los[0][0] == seg[0] and los[-1][1] == seg[1] and all(los[i][1] == los[i + 1][0] for i in range(world - 1))
            assert all(hi - lo <= seg[2] for lo, hi in los)
            used[seg[3]:seg[3] + seg[2]] += 1
        for r in range(world):                      # 1. pushes
            for p in range(world):
                if p != r:
                    for seg in segs:
                        lo, hi = shard_range(seg, p)
                        land[p][r, seg[3]:seg[3] + hi - lo] = buf[r][lo:hi]
        for r in range(world):                      # 2. the owner's sum
            for seg in segs:
                lo, hi = shard_range(seg, r)
                for s in range(world):
                    if s != r:
                        buf[r][lo:hi] += land[r][s, seg[3]:seg[3] + hi - lo]
        for r in range(world):                      # 3. pushes back
            for p in range(world):
                if p != r:
                    for seg in segs:
                        lo, hi = shard_range(seg, r)
                        buf[p][lo:hi] = buf[r][lo:hi]
    assert used.max() == 1
    total = np.sum(src, axis=0)
    exchanged = np.zeros(n, dtype=bool)
    for segs in buckets:
        for a, b in segs:
            exchanged[a:b] = True
    for r in range(world):
        np.testing.assert_allclose(buf[r][exchanged], total[exchanged], rtol=1e-12, atol=1e-12)
        np.testing.assert_array_equal(buf[r][~exchanged], src[r][~exchanged])
