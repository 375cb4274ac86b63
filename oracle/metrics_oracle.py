"""Oracle (TEST INFRASTRUCTURE): the multi-hypothesis evaluation metrics of reference ``hand/criteria.py:MHEntLoss`` and the
top-k hypothesis selection of ``hand/network.py:866-871``, restated in plain PyTorch (CPU, fp32 or fp64).

``MHEntLoss.forward`` (``criteria.py:47-173``) with ``aligned = False`` (``:62``, the shipped setting) computes, from N hypotheses per
image: per-joint 3D errors at the original scale (``utils.py:21-30``), per-joint 2D errors in pixels (``:97-105``), three joint
groups — all / visible / occluded, root joint 12 excluded from the latter two (``:107-114``) — and for each (space, group) the
best-hypothesis mean error (worst hypothesis for 2D-visible, ``:146-150``), the spread of the hypotheses per joint (``:153-159``) and,
for the visible group, the mean over hypotheses (``:161-165``).  ``_group_stats`` (``:116-131``) rescales by B / (#images that have
at least one joint in the group).
"""
from __future__ import annotations

import torch

ROOT_IDX = 12          # criteria.py:112
IMAGE_SIZE = 256.0     # criteria.py:97

METRIC_KEYS = [f'eucLoss_{sp}_rgb_{attr}{suf}' for sp in ('3d', '2d') for attr, sufs in
               (('sample', ('', '_std')), ('vis', ('', '_std', '_mean')), ('invis', ('', '_std'))) for suf in sufs]


def group_stats(stats, weight, B):
    """``_group_stats`` — reference ``criteria.py:116-131``.  stats, weight: ((N,) B, K) -> ((N,) B)."""
    num_vis = weight.sum(-1)
    mpj = (stats * weight).sum(-1) / (num_vis + 1e-16)
    if num_vis.dim() == 2:
        num_vis = num_vis[0]
    num_valid = int((num_vis > 0.).sum())
    return mpj * B / (num_valid + 1e-16) if num_valid else mpj * 0.


def hypothesis_metrics(xyz, uv, pose3d, scale, crop_uv, vis):
    """xyz (N,B,63) root-relative / bone-normalised joints, uv (N,B,42) pixels; targets pose3d (B,63), scale (B,),
    crop_uv (B,42) in [-1,1], vis (B,21).  Returns {key: (B,)} with the reference's keys (``criteria.py:91-165``)."""
    N, B = xyz.shape[:2]
    K = 21
    # criteria.py:91-95 with utils.meanEuclideanLoss(reduction='none') (utils.py:21-30)
    pred = xyz.reshape(N * B, K, 3)
    gt = pose3d.repeat(N, 1).reshape(N * B, K, 3)
    euc3 = (torch.sqrt(((pred - gt) ** 2).sum(2)) * scale.repeat(N).view(-1, 1)).reshape(N, B, K)
    uv_gt = (crop_uv + 1.) / 2. * IMAGE_SIZE                                        # :97
    euc2 = (uv - uv_gt).reshape(N, B, K, 2).norm(p=2, dim=-1)                        # :105
    weights = {'sample': torch.ones_like(vis), 'vis': (vis == 1.).to(vis.dtype), 'invis': (vis != 1.).to(vis.dtype)}   # :107-111
    weights['vis'][:, ROOT_IDX] = 0.                                                 # :113
    weights['invis'][:, ROOT_IDX] = 0.
    out = {}
    for sp, euc, D in (('3d', euc3, 3), ('2d', euc2, 2)):
        coord = (xyz.reshape(N, B, K, 3) * scale[None, :, None, None]) if sp == '3d' else uv.reshape(N, B, K, 2)   # :140-143
        for attr, w in weights.items():
            key = f'eucLoss_{sp}_rgb_{attr}'
            mpjpe = group_stats(euc, w[None].repeat(N, 1, 1), B)                     # :147
            out[key] = mpjpe.max(0)[0] if (sp == '2d' and attr == 'vis') else mpjpe.min(0)[0]   # :148-152
            spspe = torch.zeros(B, K, dtype=xyz.dtype) if N == 1 else coord.std(0).prod(-1)     # :155-159
            spspe = spspe ** (1 / D) * (D ** 0.5)                                    # :160
            out[f'{key}_std'] = group_stats(spspe, w, B)
            if attr == 'vis':
                out[f'{key}_mean'] = group_stats(euc.mean(0), w, B)                  # :163-167
    return out


def topk_hypotheses(log_q, k):
    """``network.py:866-871``: indices (k, B) of the k most likely hypotheses of every image, most likely first."""
    return torch.topk(log_q, k, dim=0)[1]
