"""Data parallelism over images: one process per GPU, weights replicated, rank g owns images
[g*B/G, (g+1)*B/G) with all their S hypotheses, so the per-image N-means and the hoisted conditioning stay
local.  The only exchange per training step is one all-reduce of the flat fp32 gradient and of the scalar loss.
The reference has no distributed code; this is the B200 addition (NCCL over NVLink; gloo in the CPU tests).
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
