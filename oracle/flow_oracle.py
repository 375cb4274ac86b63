"""Oracle (TEST INFRASTRUCTURE): conditional RealNVP of reference ``hand/flows.py``, restated.

Weights travel as a reference-compatible ``state_dict``: ``{t,s}.{i}.l.{0,1,2}.{weight,bias}``,
``{t,s}.{i}.c.{0,1}.{weight,bias}`` and ``mask`` (reference ``flows.py:83-93,189-195``).
Everything is written with explicit matmuls so the same code runs in fp32 and fp64.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

LEAKY_SLOPE = 0.01  # F.leaky_relu default, reference flows.py:117


def make_mask(dim: int, num_steps: int) -> torch.Tensor:
    """Alternating half masks, reference ``flows.py:152-155``."""
    a = [0.0] * (dim // 2) + [1.0] * (dim - dim // 2)
    b = [1.0 - v for v in a]
    return torch.tensor([a, b] * num_steps, dtype=torch.float32)


def init_state_dict(dim=45, cond_dim=512, h_dims=(512, 512), num_steps=6, seed=0, dtype=torch.float32):
    """Default ``nn.Linear`` initialisation in the reference's construction order.

    The reference builds the ``t`` ModuleList first and the ``s`` list second, each net creating
    ``l.0, l.1, l.2`` then ``c.0, c.1`` (``flows.py:83-93,190-195``); ``RealNVP._init`` is never
    called.  Drawing in the same order from the same seed reproduces the
    reference's weights bit for bit.
    """
    sd = {}

    def linear(prefix, fan_in, fan_out):
        lin = torch.nn.Linear(fan_in, fan_out)  # same init call the reference makes
        sd[prefix + '.weight'] = lin.weight.detach().to(dtype)
        sd[prefix + '.bias'] = lin.bias.detach().to(dtype)

    n_layers = 2 * num_steps
    with torch.random.fork_rng(devices=[]):
        torch.manual_seed(seed)
        for name in ('t', 's'):
            for i in range(n_layers):
                linear(f'{name}.{i}.l.0', dim, h_dims[0])
                linear(f'{name}.{i}.l.1', h_dims[0], h_dims[1])
                linear(f'{name}.{i}.l.2', h_dims[1], dim)
                linear(f'{name}.{i}.c.0', cond_dim, h_dims[0])
                linear(f'{name}.{i}.c.1', cond_dim, h_dims[1])
    sd['mask'] = make_mask(dim, num_steps).to(dtype)
    return sd


def num_layers(sd) -> int:
    return sd['mask'].shape[0]


def net_forward(sd, name: str, i: int, x, cond):
    """``_nets.forward`` — reference ``flows.py:95-122``: l2(lrelu(l1(lrelu(l0 x + c0 cond)) + c1 cond))."""
    p = f'{name}.{i}.'
    h = F.linear(x, sd[p + 'l.0.weight'], sd[p + 'l.0.bias'])
    c0 = F.linear(cond, sd[p + 'c.0.weight'], sd[p + 'c.0.bias'])
    c1 = F.linear(cond, sd[p + 'c.1.weight'], sd[p + 'c.1.bias'])
    h = F.leaky_relu(h + c0, LEAKY_SLOPE)
    h = F.linear(h, sd[p + 'l.1.weight'], sd[p + 'l.1.bias'])
    h = F.leaky_relu(h + c1, LEAKY_SLOPE)
    h = F.linear(h, sd[p + 'l.2.weight'], sd[p + 'l.2.bias'])
    if name == 's':
        h = torch.tanh(h)
    return h


def forward_p(sd, z, cond, return_logdet=False):
    """z -> x through layers 0..L-1, reference ``flows.py:210-217``.

    ``return_logdet`` additionally tracks ``sum_i sum_d s`` (the log|det dx/dz| the reference
    never materialises in this direction; SURVEY.md §4 identity).
    """
    mask = sd['mask']
    x = z
    logdet = z.new_zeros(z.shape[0])
    for i in range(num_layers(sd)):
        m = mask[i]
        x_ = x * m
        s = net_forward(sd, 's', i, x_, cond) * (1 - m)
        t = net_forward(sd, 't', i, x_, cond) * (1 - m)
        x = x_ + (1 - m) * (x * torch.exp(s) + t)
        logdet = logdet + s.sum(dim=1)
    if return_logdet:
        return x, logdet
    return x


def backward_p(sd, x, cond):
    """x -> z through layers L-1..0 with log|det dz/dx|, reference ``flows.py:219-227``."""
    mask = sd['mask']
    z = x
    log_det = x.new_zeros(x.shape[0])
    for i in reversed(range(num_layers(sd))):
        m = mask[i]
        z_ = m * z
        s = net_forward(sd, 's', i, z_, cond) * (1 - m)
        t = net_forward(sd, 't', i, z_, cond) * (1 - m)
        z = (1 - m) * (z - t) * torch.exp(-s) + z_
        log_det = log_det - s.sum(dim=1)
    return z, log_det


def std_normal_log_prob(z):
    """``MultivariateNormal(0, I).log_prob`` in closed form (reference ``flows.py:156-157,320``)."""
    d = z.shape[-1]
    return -0.5 * (z * z).sum(-1) - 0.5 * d * math.log(2 * math.pi)


def log_prob(sd, x, feat, scale=1.0, return_z=False):
    """``RealNVP.log_prob(x, logvar=feat)`` for int ``tsfm_on`` — reference ``flows.py:271-331``."""
    x = x / scale                                   # flows.py:307
    cond = feat.reshape(feat.shape[0], -1)          # make_cond, flows.py:258-268 (identity for dim=45)
    z, logdet = backward_p(sd, x, cond)             # flows.py:314
    lp = std_normal_log_prob(z) + logdet            # flows.py:320 (logdet_sigma = 0, weights = 1)
    if return_z:
        return z, lp
    return lp


def sample(sd, z0, feat, scale=1.0):
    """``RealNVP.sample`` with the prior draw ``z0 = prior.sample * temp`` supplied by the caller.

    Reference ``flows.py:333-359``; the draw itself (``flows.py:339``) equals ``torch.randn``
    under the same seed and stays outside the oracle so all arms share it.
    """
    cond = feat.reshape(feat.shape[0], -1)
    x = forward_p(sd, z0, cond)
    return x * scale


def cast_state_dict(sd, dtype):
    return {k: v.detach().to(dtype) for k, v in sd.items()}
