#!/usr/bin/env python
"""Precision-scheme evidence: error of the flow's tensor-core arithmetic candidates against fp64.

CPU emulation (test infrastructure: it drives the ORACLE, not the product).  Every ``F.linear`` of the flow
(coupling ``l.{0,1,2}`` and conditioning ``c.{0,1}``, reference ``hand/flows.py:105-122``) is replaced by an
autograd function that rounds its operands the way a tensor-core scheme would and accumulates exactly (fp64):

  fp32        plain fp32 matmuls (the reference's own arithmetic; its distance to fp64 is the FLOOR)
  tf32x1      both operands rounded to 11 significant bits (one ``kind::tf32`` pass; ``tf32x1_trunc`` = truncated)
  3p          x = hi + lo 16-bit planes on both sides, hi*hi + hi*lo + lo*hi   (round 1's scheme: 3 tensor passes)
  2p_w1       weights as ONE plane (hi), activations / gradients as two:  W_hi * (x_hi + x_lo)   (2 passes)
  2p_a1       weights as two planes, activations / gradients as one       (2 passes)
  1p          one plane each (1 pass)

``fwd`` planes are IEEE half (11 bits each) as in the kernels; ``bwd`` planes are bfloat16 (8 bits each) unless the
scheme name ends in ``_h`` (half planes for the backward's weight / activation operand, bfloat16 for the gradient only -
tcgen05 ``kind::f16`` takes the A and B formats separately).

Writes ``profiles/r2_precision_schemes.json``: relative errors of log q, x, the loss, the flat parameter gradient
(Frobenius), dfeat and dz_det on BASELINE configs[1]'s shape (B=64 x S=10) against the fp64 oracle.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import types

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import flow_oracle as fo, loss_oracle as lo, mano_oracle as mo  # noqa: E402
from mhentropy_b200.mano_assets import synthetic_mano  # noqa: E402


def r_half(x):
    return x.to(torch.float16).to(x.dtype)


def r_bf16(x):
    return x.to(torch.bfloat16).to(x.dtype)


def r_tf32(x, trunc=False):
    i = x.to(torch.float32).contiguous().view(torch.int32)
    if trunc:
        i = i & ~0x1FFF
    else:
        i = (i + 0xFFF + ((i >> 13) & 1)) & ~0x1FFF
    return i.view(torch.float32).to(x.dtype)


def planes(x, r, n):
    """x as n planes under rounding r: returns (hi, lo) with lo = 0 when n == 1."""
    hi = r(x)
    if n == 1:
        return hi, torch.zeros_like(hi)
    return hi, r(x - hi)


def product(a, b, ra, rb, na, nb):
    """a @ b.T with a as na planes (rounding ra), b as nb planes (rounding rb); the lo*lo term is never issued."""
    a = a.double()
    b = b.double()
    ah, al = planes(a, ra, na)
    bh, bl = planes(b, rb, nb)
    out = ah @ bh.T
    if nb == 2:
        out = out + ah @ bl.T
    if na == 2:
        out = out + al @ bh.T
    return out


class SchemeLinear(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b, sch):
        ctx.save_for_backward(x, w)
        ctx.sch = sch
        y = product(x, w, sch['f_act'], sch['f_w'], sch['n_act'], sch['n_w'])
        return (y + b.double()).to(x.dtype)

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        s = ctx.sch
        # dx = dy @ w: gradient (activation-side operand) against the weights
        dx = product(dy, w.T, s['b_grad'], s['b_w'], s['bn_act'], s['bn_w']).to(x.dtype)
        # dw = dy.T @ x: gradient against the saved activation (the weight-side role is played by the saved activation)
        dw = product(dy.T, x.T, s['b_grad'], s['b_act'], s['bn_act'], s['bn_w']).to(w.dtype)
        return dx, dw, dy.sum(0).to(dy.dtype), None


def scheme(name):
    ident = lambda x: x  # noqa: E731
    if name == 'fp32':
        return None
    if name.startswith('tf32x1'):
        r = (lambda x: r_tf32(x, True)) if name.endswith('trunc') else r_tf32
        return dict(f_act=r, f_w=r, b_grad=r, b_w=r, b_act=r, n_act=1, n_w=1, bn_act=1, bn_w=1)
    del ident
    counts = {'3p': (2, 2), '2p_w1': (2, 1), '2p_a1': (1, 2), '1p': (1, 1)}
    if '+' in name:                       # 'F<scheme>+B<scheme>[_h]': forward and backward chosen separately
        fname, bname = name.split('+')
        fname, bname = fname[1:], bname[1:]
    else:
        fname = bname = name
    bbase, half_bwd = (bname[:-2], True) if bname.endswith('_h') else (bname, False)
    fbase = fname[:-2] if fname.endswith('_h') else fname
    bw = r_half if half_bwd else r_bf16
    fa, fw = counts[fbase]
    ba, bwn = counts[bbase]
    return dict(f_act=r_half, f_w=r_half, b_grad=r_bf16, b_w=bw, b_act=bw, n_act=fa, n_w=fw, bn_act=ba, bn_w=bwn)


def run(sd, c, batch, S, sch, dtype):
    sdg = {k: v.detach().to(dtype).requires_grad_(k != 'mask') for k, v in sd.items()}
    feat = batch['feat'].to(dtype).clone().requires_grad_(True)
    zd = batch['z_det'].to(dtype).clone().requires_grad_(True)
    z0 = batch['z0'].to(dtype)
    shim = types.SimpleNamespace(**{k: getattr(F, k) for k in ('leaky_relu', 'relu')})
    shim.linear = F.linear if sch is None else (lambda x, w, b: SchemeLinear.apply(x, w, b, sch))
    old = fo.F
    fo.F = shim
    try:
        x, logdet = fo.forward_p(sdg, z0, feat.repeat(S, 1), return_logdet=True)
    finally:
        fo.F = old
    log_q = fo.std_normal_log_prob(z0) - logdet
    z = lo.combine_z(x, zd.repeat(S, 1))
    cm = {k: (v.to(dtype) if torch.is_tensor(v) and v.is_floating_point() else v) for k, v in c.items()}
    terms = lo.forward_log_p(cm, z, batch['crop_uv'].to(dtype), batch['vis'].to(dtype), S)
    log_p = (-log_q).reshape(S, -1).mean(0) + terms['log_p'].reshape(S, -1).mean(0)
    loss = lo.mhent_loss(log_p)
    loss.backward()
    names = [k for k in sdg if k != 'mask']
    flat = torch.cat([sdg[k].grad.flatten().double() for k in names])
    return {'log_q': log_q.detach().double(), 'x': x.detach().double(), 'loss': loss.detach().double().reshape(1), 'flat': flat,
            'dfeat': feat.grad.double(), 'dz_det': zd.grad.double(),
            'per_tensor': {k: sdg[k].grad.double() for k in names}}


def rel(a, b):
    return float((a - b).norm() / b.norm())


def relmax(a, b):
    return float((a - b).abs().max() / b.abs().max())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--batch', type=int, default=64)
    ap.add_argument('--hyp', type=int, default=10)
    ap.add_argument('--schemes', default='fp32,tf32x1,tf32x1_trunc,3p,3p_h,2p_w1,2p_w1_h,2p_a1,2p_a1_h,1p_h,F3p+B2p_w1,F3p+B2p_w1_h,F3p+B2p_a1,F3p+B2p_a1_h,F2p_w1+B3p')
    ap.add_argument('--out', default=os.path.join(ROOT, 'profiles', 'r2_precision_schemes.json'))
    args = ap.parse_args()
    torch.set_num_threads(os.cpu_count() or 1)
    sd = fo.init_state_dict(seed=0)
    c = mo.mano_constants(synthetic_mano(0))
    batch = lo.synthetic_batch(args.batch, args.hyp, seed=0)
    ref = run(sd, c, batch, args.hyp, None, torch.float64)
    rows = {}
    for name in args.schemes.split(','):
        got = run(sd, c, batch, args.hyp, scheme(name), torch.float32)
        worst = sorted(((rel(got['per_tensor'][k], ref['per_tensor'][k]), k) for k in ref['per_tensor']), reverse=True)
        rows[name] = {
            'log_q_relmax': relmax(got['log_q'], ref['log_q']), 'x_relmax': relmax(got['x'], ref['x']),
            'loss_rel': rel(got['loss'], ref['loss']), 'flat_grad_rel_fro': rel(got['flat'], ref['flat']),
            'dfeat_relmax': relmax(got['dfeat'], ref['dfeat']), 'dfeat_rel_fro': rel(got['dfeat'], ref['dfeat']),
            'dz_det_relmax': relmax(got['dz_det'], ref['dz_det']),
            'param_tensors_above_1e-3': sum(1 for e, _ in worst if e > 1e-3), 'worst_param_tensor': [worst[0][1], worst[0][0]],
        }
        print(name, json.dumps(rows[name]), flush=True)
    out = {'what': __doc__.split('\n')[0], 'shape': {'B': args.batch, 'S': args.hyp}, 'bars': {'log_prob': 1e-4, 'gradients': 1e-3},
           'schemes': rows}
    with open(args.out, 'w') as fh:
        json.dump(out, fh, indent=1)


if __name__ == '__main__':
    main()
