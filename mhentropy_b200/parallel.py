"""Data parallelism over images: one process per GPU, weights replicated, rank g owns images
[g*B/G, (g+1)*B/G) with all their S hypotheses, so the per-image N-means and the hoisted conditioning stay
local.  The only exchange per training step is one all-reduce of the flat fp32 gradient and of the scalar loss.
The reference has no distributed code; this is the B200 addition (NCCL over NVLink; gloo in the CPU tests).

:class:`PeerExchange` is the exchange that overlaps with the step: the gradient lives in symmetric (peer-mapped) memory, shards travel
by copy-engine pushes over NVLink / NVSwitch (no SM is taken from the step's cluster kernels, unlike NCCL's), and one small kernel of
this library (``mhe_sum_shards``) adds the landed copies.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def image_range(B: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous, balanced image range of ``rank`` (first ``B % world`` ranks get one extra image)."""
    base, rem = divmod(B, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(batch: dict, S: int, rank: int, world: int) -> dict:
    """Slice a global batch by image.  ``z0`` is hypothesis-major (row r = n*B + b, reference network.py:734),
    so its shard keeps every hypothesis n of the local images, again hypothesis-major."""
    B = batch['feat'].shape[0]
    lo, hi = image_range(B, rank, world)
    out = {}
    for k, v in batch.items():
        if k == 'z0':
            out[k] = v.reshape(S, B, -1)[:, lo:hi].reshape(S * (hi - lo), -1).contiguous()
        else:
            out[k] = v[lo:hi].contiguous()
    return out


def allreduce_step(flat_grad: torch.Tensor, loss: torch.Tensor, local_images: int, global_images: int, group=None):
    """Turn per-rank (mean-over-local-images) loss and gradient into the global-batch mean, in place.

    loss_global = sum_g (B_g / B) loss_g, likewise for the gradient: scale locally, then sum-all-reduce.
    """
    w = local_images / global_images
    flat_grad.mul_(w)
    loss.mul_(w)
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat_grad, group=group)
        dist.all_reduce(loss, group=group)
    return flat_grad, loss


def gather_cond_factors(dcp: torch.Tensor, feat: torch.Tensor, group=None):
    """Factored exchange of the conditioning weight gradient: ``dCw[idx] = dcp[:, idx, :]^T feat`` has rank <= images
    per weight matrix, so the ranks all-gather the factors - ``dcp`` (B, L*4*H) and ``feat`` (B, C), equal B on every rank - instead of
    all-reducing the dense gradient.  Returns ``(dcp_all, feat_all)`` with world*B rows; ``mhe_flow_cond_wgrad`` (or
    :func:`cond_wgrad_from_factors` on the CPU) turns them into the global gradient."""
    if not (dist.is_initialized() and dist.get_world_size(group) > 1):
        return dcp, feat
    world = dist.get_world_size(group)
    dcp_all = dcp.new_empty(world * dcp.shape[0], dcp.shape[1])
    feat_all = feat.new_empty(world * feat.shape[0], feat.shape[1])
    dist.all_gather_into_tensor(dcp_all, dcp.contiguous(), group=group)
    dist.all_gather_into_tensor(feat_all, feat.contiguous(), group=group)
    return dcp_all, feat_all


def cond_wgrad_from_factors(dcp_all: torch.Tensor, feat_all: torch.Tensor, hidden: int) -> torch.Tensor:
    """Plain-PyTorch statement of ``mhe_flow_cond_wgrad``: (L*4, H, C) conditioning weight gradients from the (gathered) factors."""
    Bt = dcp_all.shape[0]
    return torch.einsum('bih,bc->ihc', dcp_all.reshape(Bt, -1, hidden), feat_all)


# ---------------------------------------------------------------------------------------------------------------------------------------
# peer-memory exchange: reduce-scatter + all-gather by copy-engine pushes
def shard_length(n: int, world: int, align: int = 32) -> int:
    """Floats per rank of an ``n``-float segment cut into ``world`` contiguous shards (a multiple of ``align``; trailing shards may be
    short or empty)."""
    return -(-(-(-n // world)) // align) * align


def plan_peer_buckets(buckets, world: int, align: int = 32):
    """``buckets``: list of lists of (start, stop) float ranges of the flat buffer.  Returns ``(plan, landing_floats)``: per bucket a list
    of (start, stop, shard, landing_offset); rank p owns [start + p*shard, min(stop, start + (p+1)*shard)) of every segment and receives
    the other ranks' copies of it at ``landing_offset`` of its landing slot for the source rank."""
    plan, off = [], 0
    for segs in buckets:
        out = []
        for a, b in segs:
            if b < a:
                raise ValueError('empty-negative segment')
            sh = shard_length(b - a, world, align)
            out.append((a, b, sh, off))
            off += sh
        plan.append(out)
    return plan, off


def shard_range(seg, rank: int):
    """[lo, hi) of ``rank``'s shard of a planned segment (possibly empty)."""
    a, b, sh, _ = seg
    lo = min(b, a + rank * sh)
    return lo, min(b, lo + sh)


class PeerExchange:
    """Sum-all-reduce of ranges of a flat fp32 buffer over peer-mapped memory.  Collective constructor: allocates ``buf`` (``n_floats``
    rounded up + ``extra`` floats for scalars such as the loss) in symmetric memory.  ``plan(buckets)`` (collective) fixes the ranges
    exchanged together; ``reduce_bucket(i)`` enqueues, on the current stream and capturable in a CUDA graph:

      1. per peer p, on its own stream: copy of my values of p's shard into p's landing slot for me (cudaMemcpyAsync between peer-mapped
         allocations = a copy engine; writes are posted, so pushes beat pulls on NVLink),
      2. a device-side barrier over the ranks, ``mhe_sum_shards`` on my shard (+= the world-1 landed copies),
      3. per peer: copy of my reduced shard into p's buffer, and a closing barrier.

    Hazards: a landing slot is rewritten only after its owner passed the closing barrier of the previous use; a peer's buffer range is
    overwritten in (3) only after that peer pushed its copy of it (it reached the barrier of (2)).  The caller must not write the
    exchanged ranges again before ``reduce_bucket`` has completed in stream order."""

    ALIGN = 32

    def __init__(self, n_floats: int, device, group=None, extra: int = 32):
        import os
        import torch.distributed._symmetric_memory as symm_mem
        self._symm_mem = symm_mem
        self.group = group if group is not None else dist.group.WORLD
        self.rank, self.world = dist.get_rank(self.group), dist.get_world_size(self.group)
        self.dev = torch.device(device)
        self.n = int(n_floats)
        self.n_pad = -(-self.n // self.ALIGN) * self.ALIGN
        self.size = self.n_pad + int(extra)
        self.buf = symm_mem.empty(self.size, dtype=torch.float32, device=self.dev)
        self.buf.zero_()
        self.hdl = symm_mem.rendezvous(self.buf, self.group.group_name)
        self.peers = [(self.rank + i) % self.world for i in range(1, self.world)]
        self.peer_buf = {p: self.hdl.get_buffer(p, (self.size,), torch.float32) for p in self.peers}
        k = int(os.environ.get("MHE_PEER_COPIES", 1))      # >1: measured slower (parallel copies to ONE peer contend)
        self.streams = [[torch.cuda.Stream(self.dev) for _ in range(k)] for _ in self.peers]
        self.buckets, self.land = None, None
        self.mark = None            # optional callable(label): phase marks for a timeline (engine.TrainStep.trace)

    def _mark(self, label):
        if self.mark is not None:
            self.mark(label)

    def plan(self, buckets):
        self.buckets, self.land_floats = plan_peer_buckets(buckets, self.world, self.ALIGN)
        self.land_floats = max(self.land_floats, self.ALIGN)
        self.land = self._symm_mem.empty(self.world * self.land_floats, dtype=torch.float32, device=self.dev)
        self.hland = self._symm_mem.rendezvous(self.land, self.group.group_name)
        self.peer_land = {p: self.hland.get_buffer(p, (self.world, self.land_floats), torch.float32) for p in self.peers}
        torch.cuda.synchronize(self.dev)
        dist.barrier(self.group)
        return self

    def _fan_out(self, copies_for):
        """Enqueue the (dst, src) copies of every peer on that peer's streams, forked from and joined back into the current stream.  One
        copy engine moves ~450 GB/s over NVLink; with few peers a large copy is cut over ``copies_per_peer`` streams to fill the link."""
        main = torch.cuda.current_stream(self.dev)
        used = []
        for sts, p in zip(self.streams, self.peers):
            jobs = []
            for dst, src in copies_for(p):
                k = len(sts) if src.numel() >= (1 << 20) else 1
                step = -(-src.numel() // k)
                jobs += [(dst[j:j + step], src[j:j + step]) for j in range(0, src.numel(), step)]
            for j, (dst, src) in enumerate(jobs):
                st = sts[j % len(sts)]
                if st not in used:
                    st.wait_stream(main)
                    used.append(st)
                with torch.cuda.stream(st):
                    dst.copy_(src, non_blocking=True)
        for st in used:
            main.wait_stream(st)

    def reduce_bucket(self, i: int, closing: bool = True):
        """``closing=False`` leaves out the closing barrier: allowed when another ``reduce_bucket`` follows in stream order before anything
        reads the exchanged ranges or rewrites them (its barriers order this bucket's pushes as well)."""
        from . import _lib
        segs, r = self.buckets[i], self.rank

        def scatter(p):
            out = []
            for seg in segs:
                lo, hi = shard_range(seg, p)
                if hi > lo:
                    out.append((self.peer_land[p][r, seg[3]:seg[3] + hi - lo], self.buf[lo:hi]))
            return out

        self._mark(f'bucket{i}.start')
        self._fan_out(scatter)
        self._mark(f'bucket{i}.scattered')
        self.hdl.barrier(channel=0)
        self._mark(f'bucket{i}.barrier0')
        L, sp = _lib.lib(), _lib.stream_ptr(self.dev)
        mine = []
        for seg in segs:
            lo, hi = shard_range(seg, r)
            if hi > lo:
                _lib.check(L.mhe_sum_shards(self.buf.data_ptr() + 4 * lo, self.land.data_ptr() + 4 * seg[3], self.world, r,
                                            self.land_floats, hi - lo, sp), 'sum_shards')
                mine.append((lo, hi))
        self._mark(f'bucket{i}.summed')
        self._fan_out(lambda p: [(self.peer_buf[p][lo:hi], self.buf[lo:hi]) for lo, hi in mine])
        self._mark(f'bucket{i}.gathered')
        if closing:
            self.hdl.barrier(channel=1)
            self._mark(f'bucket{i}.barrier1')

    def all_reduce(self):
        """Every planned bucket, in order."""
        for i in range(len(self.buckets)):
            self.reduce_bucket(i, closing=i == len(self.buckets) - 1)
