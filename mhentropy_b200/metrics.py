"""Hypothesis selection and the multi-hypothesis evaluation metrics on the B200 kernels.

* :func:`topk_hypotheses`  <- ``torch.topk(log_q, N_quant, dim=0)[1]`` in reference ``hand/network.py:866-871``
* :class:`MHEntLoss`       <- reference ``hand/criteria.py:40-173`` (``aligned = False``, the shipped setting): same call
  signature and return triple ``(loss, losses, metrics)``, same metric keys.

One block per image reads the (N, B, .) outputs of ``MHEnt.sample`` exactly once; nothing else is materialised.
CUDA tensors always take the kernels (``MheError`` if the library is missing) — no CPU fallback.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _lib
from ._lib import check, lib, ptr, stream_ptr

ROOT_IDX = 12        # criteria.py:112
IMAGE_SIZE = 256.0   # criteria.py:97

METRIC_KEYS = [f'eucLoss_{sp}_rgb_{attr}{suf}' for sp in ('3d', '2d') for attr, sufs in
               (('sample', ('', '_std')), ('vis', ('', '_std', '_mean')), ('invis', ('', '_std'))) for suf in sufs]


def topk_hypotheses(log_q: torch.Tensor, k: int) -> torch.Tensor:
    """log_q (N, B) -> int64 indices (k, B) of the k most likely hypotheses of every image, most likely first."""
    N, B = log_q.shape
    if not 1 <= k <= N:
        raise ValueError(f'topk_hypotheses: k = {k} out of range for N = {N}')
    log_q = log_q.contiguous().float()
    _lib.require_cuda_f32(log_q)
    idx = torch.empty(k, B, dtype=torch.int64, device=log_q.device)
    check(lib().mhe_topk_hypotheses(ptr(log_q), N, B, k, ptr(idx), stream_ptr(log_q.device)), 'mhe_topk_hypotheses')
    return idx


def hypothesis_metrics(xyz, uv, pose3d, scale, crop_uv, vis) -> dict:
    """xyz (N,B,63), uv (N,B,42) pixels; pose3d (B,63), scale (B,), crop_uv (B,42), vis (B,21) -> {key: (B,)}."""
    N, B = xyz.shape[:2]
    args = [t.contiguous().float() for t in (xyz, uv, pose3d, scale, crop_uv, vis)]
    _lib.require_cuda_f32(*args)
    dev = args[0].device
    out = torch.empty(len(METRIC_KEYS), B, device=dev)
    nbytes = lib().mhe_hypothesis_metrics_workspace_bytes(B)
    ws = torch.empty(max(nbytes, 4), dtype=torch.uint8, device=dev)
    check(lib().mhe_hypothesis_metrics(*[ptr(t) for t in args], N, B, ROOT_IDX, IMAGE_SIZE, ptr(out), ptr(ws), nbytes,
                                       stream_ptr(dev)), 'mhe_hypothesis_metrics')
    return {k: out[i] for i, k in enumerate(METRIC_KEYS)}


class MHEntLoss(nn.Module):
    """Drop-in for reference ``criteria.py:MHEntLoss``: ``forward(output, target) -> (loss, losses, metrics)``."""

    def __init__(self, loss_weights=None):
        super().__init__()
        self.loss_weights = loss_weights

    def forward(self, output: dict, target: dict):
        losses = {'neg_log_p': -output['log_p']}                       # criteria.py:53
        if 'uv' in output:
            uv = output['uv']
        else:                                                          # criteria.py:98-104: GT scale / translation
            xyz_ = output['xyz'].reshape(*output['xyz'].shape[:-1], -1, 3)
            uv = target['st'][:, None, [0]] * xyz_[..., :2] + target['st'][:, None, -2:]
            output['uv'] = uv = ((uv + 1) / 2 * IMAGE_SIZE).flatten(start_dim=-2)
        metrics = hypothesis_metrics(output['xyz'], uv, target['pose3d'], target['scale'], target['crop_uv'], target['vis'])
        return sum(v.mean() for v in losses.values()), losses, metrics  # criteria.py:173
