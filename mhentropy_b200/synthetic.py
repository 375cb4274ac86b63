"""Synthetic inputs of the measured path: same recipe for the CUDA arm, the oracle and the tests."""
from __future__ import annotations

import math

import torch


def synthetic_batch(B, S, seed=0, dtype=torch.float32, temp=1.0, cond_dim=512, dim=45):
    """feat ~ N(0,1) (B,cond); z_det = th3 0.5N | beta 0.02N | logs N(log .3, .1) | t 0.1N (B,16);
    z0 = randn (S*B, dim) * temp (== prior.sample * temp, reference flows.py:339); crop_uv ~ U(-1,1) (B,42);
    vis ~ Bernoulli(0.7) (B,21).  Drawn in this order from one CPU generator seeded with ``seed``."""
    g = torch.Generator().manual_seed(seed)
    feat = torch.randn(B, cond_dim, generator=g)
    th3 = 0.5 * torch.randn(B, 3, generator=g)
    beta = 0.02 * torch.randn(B, 10, generator=g)
    logs = math.log(0.3) + 0.1 * torch.randn(B, 1, generator=g)
    t = 0.1 * torch.randn(B, 2, generator=g)
    z0 = torch.randn(B * S, dim, generator=g) * temp
    crop_uv = torch.rand(B, 42, generator=g) * 2 - 1
    vis = (torch.rand(B, 21, generator=g) < 0.7).float()
    z_det = torch.cat([th3, beta, logs, t], dim=1)
    return {'feat': feat.to(dtype), 'z_det': z_det.to(dtype), 'z0': z0.to(dtype),
            'crop_uv': crop_uv.to(dtype), 'vis': vis.to(dtype)}


def det_head_state_dict(seed=5, feat_dim=512):
    """Seeded weights for ``det_head`` (reference ``network.py:376-385``: Linear(512, 512) - ReLU - Linear(512, 16)) whose outputs land in the
    range of a trained head: th3 ~ 0.5, beta ~ 0.02, log s ~ log 0.3, t ~ 0.1.  Shared by the golden-fixture generator (which loads them into
    the REFERENCE's det_head) and the tests (which load them into ``MHEntHead.det_head``), so the fixture need not carry a megabyte of weights."""
    g = torch.Generator().manual_seed(seed)
    w0 = torch.randn(feat_dim, feat_dim, generator=g) / math.sqrt(feat_dim)
    b0 = 0.1 * torch.randn(feat_dim, generator=g)
    scale = torch.cat([0.5 * torch.ones(3), 0.02 * torch.ones(10), 0.1 * torch.ones(1), 0.1 * torch.ones(2)])
    w2 = torch.randn(16, feat_dim, generator=g) / math.sqrt(feat_dim) * scale[:, None]
    b2 = torch.zeros(16)
    b2[13] = math.log(0.3)
    return {'0.weight': w0, '0.bias': b0, '2.weight': w2, '2.bias': b2}
