#!/usr/bin/env python
"""Gradient error budget of the training step against fp64 autograd of the oracle.

Separates, on the same data, the reference's OWN fp32 noise (oracle fp32 vs fp64), the exact-fp32 CUDA path and the tensor-core
(split bf16x3) CUDA path, stage by stage:

  flow     sample_with_log_prob forward + backward with random output seeds          -> dz0, dfeat, flat parameter gradient
  rows     MANO joints + reprojection + priors forward / backward of given z          -> dz      (mhe_hypothesis_rows_fwd_bwd)
  step     the whole step (TrainStep engine and the autograd drop-in path)            -> dfeat, dz_det, flat gradient, log_p

Run on a GPU box:  python tools/grad_error_budget.py [--out gpurun_out/grad_error_budget.json]
(MHE_FUSED_MAX_ROWS=0 in the environment sends the tensor-core runs through the per-GEMM path instead of the cluster-fused one).
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from mhentropy_b200 import MHEntHead, RealNVP, _lib  # noqa: E402
from mhentropy_b200._lib import check, lib, ptr  # noqa: E402
from mhentropy_b200.engine import TrainStep  # noqa: E402
from mhentropy_b200.mano_assets import synthetic_mano  # noqa: E402
from mhentropy_b200.synthetic import synthetic_batch  # noqa: E402
from oracle import flow_oracle as fo, loss_oracle as lo, mano_oracle as mo  # noqa: E402

DEV = 'cuda'
PROD = dict(dim=45, tsfm_on=512, kemb=False, jointN=21, h_dims=[512, 512], num_steps=6)


def fro(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-300))


def mx(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / (b.abs().max() + 1e-300))


def flat_err(grads: dict, ref: dict):
    num = sum(float((grads[k].detach().double().cpu() - ref[k]).pow(2).sum()) for k in ref)
    den = sum(float(ref[k].pow(2).sum()) for k in ref)
    per = sorted(((fro(grads[k], ref[k]), k) for k in ref), reverse=True)
    return {'flat_fro': (num / den) ** 0.5, 'tensors_above_1e-3': sum(1 for e, _ in per if e > 1e-3), 'worst': [per[0][1], per[0][0]]}


# ---------------------------------------------------------------------------------------------------------------------
def oracle_flow(sd, feat, z0, wx, wl, S, dtype):
    sdg = {k: v.to(dtype).clone().requires_grad_(k != 'mask') for k, v in sd.items()}
    f = feat.to(dtype).clone().requires_grad_(True)
    z = z0.to(dtype).clone().requires_grad_(True)
    x, ld = fo.forward_p(sdg, z, f.repeat(S, 1), return_logdet=True)
    logq = fo.std_normal_log_prob(z) - ld
    ((x * wx.to(dtype)).sum() + (logq * wl.to(dtype)).sum()).backward()
    return {'x': x.detach().double(), 'log_q': logq.detach().double(), 'dz0': z.grad.double(), 'dfeat': f.grad.double(),
            'params': {k: v.grad.double() for k, v in sdg.items() if k != 'mask'}}


def cuda_flow(sd, feat, z0, wx, wl, S, precision):
    flow = RealNVP(**PROD)
    flow.load_state_dict(sd)
    flow.precision = precision
    flow = flow.to(DEV)
    f = feat.to(DEV).requires_grad_(True)
    z = z0.to(DEV).requires_grad_(True)
    x, logq = flow.sample_with_log_prob(f, z, S)
    ((x * wx.to(DEV)).sum() + (logq * wl.to(DEV)).sum()).backward()
    torch.cuda.synchronize()
    return {'x': x, 'log_q': logq, 'dz0': z.grad, 'dfeat': f.grad, 'params': {k: p.grad for k, p in flow.named_parameters()}}


def stage_flow(B, S, seed):
    sd = fo.init_state_dict(seed=0)
    g = torch.Generator().manual_seed(seed)
    feat, z0 = torch.randn(B, 512, generator=g), torch.randn(B * S, 45, generator=g)
    wx, wl = torch.randn(B * S, 45, generator=g), torch.randn(B * S, generator=g)
    ref = oracle_flow(sd, feat, z0, wx, wl, S, torch.float64)
    rows = {}
    for name, got in (('oracle_fp32', oracle_flow(sd, feat, z0, wx, wl, S, torch.float32)),
                      ('cuda_fp32', cuda_flow(sd, feat, z0, wx, wl, S, 'fp32')),
                      ('cuda_bf16x3', cuda_flow(sd, feat, z0, wx, wl, S, 'bf16x3'))):
        rows[name] = {'x_absmax': float((got['x'].detach().double().cpu() - ref['x']).abs().max()), 'log_q_relmax': mx(got['log_q'], ref['log_q']),
                      'dz0_fro': fro(got['dz0'], ref['dz0']), 'dfeat_fro': fro(got['dfeat'], ref['dfeat']), 'dfeat_relmax': mx(got['dfeat'], ref['dfeat']),
                      **flat_err(got['params'], ref['params'])}
    return rows


# ---------------------------------------------------------------------------------------------------------------------
def stage_rows(B, S, seed):
    """MANO joints + reprojection + priors of given z: dz of sum_b log_p-part (dloss = 1 semantics of the kernel: seed -1/R per row)."""
    mano = synthetic_mano(0)
    R = B * S
    g = torch.Generator().manual_seed(seed)
    z = torch.cat([0.5 * torch.randn(R, 3, generator=g), 1.0 * torch.randn(R, 45, generator=g), 0.02 * torch.randn(R, 10, generator=g),
                   -1.2 + 0.1 * torch.randn(R, 1, generator=g), 0.1 * torch.randn(R, 2, generator=g)], 1).contiguous()
    crop_uv = torch.rand(B, 42, generator=g) * 2 - 1
    vis = (torch.rand(B, 21, generator=g) < 0.7).float()

    def oracle(dtype):
        c = mo.mano_constants(mano, dtype)
        zz = z.to(dtype).clone().requires_grad_(True)
        terms = lo.forward_log_p(c, zz, crop_uv.to(dtype), vis.to(dtype), S)
        (-terms['log_p'].sum() / R).backward()       # loss = mean_b(-mean_n row_log_p) = -sum_r row_log_p / R
        return {'row_log_p': terms['log_p'].detach().double(), 'dz': zz.grad.double()}

    ref = oracle(torch.float64)
    o32 = oracle(torch.float32)
    head = MHEntHead(mano_data=mano).to(DEV)
    consts, cfg, L = head.mano_dec.mano_layer._consts(torch.device(DEV)), head.loss_cfg, lib()
    zc, cuv, visc = z.to(DEV), crop_uv.to(DEV), vis.to(DEV)      # (keep references: the kernel reads them after this line)
    f = lambda *sh: torch.empty(*sh, device=DEV)  # noqa: E731
    jtr, uv, lp, dz, dlq = f(R, 21, 3), f(R, 42), f(R), f(R, 61), f(R)
    s = _lib.stream_ptr(torch.device(DEV))
    check(L.mhe_hypothesis_rows_fwd_bwd(consts, cfg, ptr(zc), None, None, ptr(cuv), ptr(visc), R, B, 1, 1.0, None, ptr(jtr), ptr(uv),
                                        ptr(lp), ptr(dz), None, ptr(dlq), s), 'rows')
    torch.cuda.synchronize()
    out = {}
    for name, got in (('oracle_fp32', o32), ('cuda', {'row_log_p': lp, 'dz': dz})):
        d, r = got['dz'].detach().double().cpu(), ref['dz']
        out[name] = {'row_log_p_relmax': mx(got['row_log_p'], ref['row_log_p']), 'dz_fro': fro(d, r), 'dz_relmax': mx(d, r),
                     'dz_theta_fro': fro(d[:, :48], r[:, :48]), 'dz_beta_fro': fro(d[:, 48:58], r[:, 48:58]), 'dz_cam_fro': fro(d[:, 58:], r[:, 58:])}
    return out


# ---------------------------------------------------------------------------------------------------------------------
def stage_step(B, S, seed):
    mano = synthetic_mano(0)
    sd = fo.init_state_dict(seed=0)
    batch = synthetic_batch(B, S, seed=seed)

    def oracle(dtype):
        sdg = {k: v.to(dtype).clone().requires_grad_(k != 'mask') for k, v in sd.items()}
        feat = batch['feat'].to(dtype).clone().requires_grad_(True)
        zd = batch['z_det'].to(dtype).clone().requires_grad_(True)
        out = lo.reverse_kld(sdg, mo.mano_constants(mano, dtype), feat, zd, batch['z0'].to(dtype), batch['crop_uv'].to(dtype), batch['vis'].to(dtype), S)
        lo.mhent_loss(out['log_p']).backward()
        return {'log_p': out['log_p'].detach().double(), 'dfeat': feat.grad.double(), 'dz_det': zd.grad.double(),
                'params': {k: v.grad.double() for k, v in sdg.items() if k != 'mask'}}

    ref = oracle(torch.float64)
    rows = {}

    def add(name, got):
        rows[name] = {'log_p_relmax': mx(got['log_p'], ref['log_p']), 'dfeat_fro': fro(got['dfeat'], ref['dfeat']), 'dfeat_relmax': mx(got['dfeat'], ref['dfeat']),
                      'dz_det_fro': fro(got['dz_det'], ref['dz_det']), 'dz_det_relmax': mx(got['dz_det'], ref['dz_det']), **flat_err(got['params'], ref['params'])}

    add('oracle_fp32', oracle(torch.float32))
    devb = {k: v.to(DEV) for k, v in batch.items()}
    for precision in ('fp32', 'bf16x3'):
        head = MHEntHead(mano_data=mano)
        head.q_z_giv_i.load_state_dict(sd)
        head.q_z_giv_i.precision = precision
        head = head.to(DEV)
        eng = TrainStep(head, B, S, DEV, want_verts=False, use_graph=False)
        eng.load(**devb)
        eng.run()
        torch.cuda.synchronize()
        add(f'engine_{precision}', {'log_p': eng.log_p, 'dfeat': eng.dfeat, 'dz_det': eng.dz_det, 'params': eng.flow_grads()})
        feat = devb['feat'].clone().requires_grad_(True)
        zd = devb['z_det'].clone().requires_grad_(True)
        for p in head.parameters():
            p.requires_grad_(True)
        head.zero_grad(set_to_none=True)
        out = head.get_loss(feat, {'crop_uv': devb['crop_uv'], 'vis': devb['vis']}, z0=devb['z0'], z_det=zd, N=S)
        (-out['log_p']).mean().backward()
        torch.cuda.synchronize()
        add(f'autograd_{precision}', {'log_p': out['log_p'], 'dfeat': feat.grad, 'dz_det': zd.grad,
                                      'params': {k: p.grad for k, p in head.q_z_giv_i.named_parameters()}})
    return rows


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--out', default=os.path.join(ROOT, 'gpurun_out', 'grad_error_budget.json'))
    ap.add_argument('--shapes', default='4x10:7,64x10:0')
    args = ap.parse_args()
    torch.set_num_threads(os.cpu_count() or 1)
    res = {'fused_max_rows_env': os.environ.get('MHE_FUSED_MAX_ROWS'), 'shapes': {}}
    for spec in args.shapes.split(','):
        shp, seed = spec.split(':')
        B, S = (int(v) for v in shp.split('x'))
        r = {'flow': stage_flow(B, S, int(seed) + 77), 'rows': stage_rows(B, S, int(seed) + 5), 'step': stage_step(B, S, int(seed))}
        res['shapes'][spec] = r
        for stage, rows in r.items():
            for name, vals in rows.items():
                print(f'{spec:10s} {stage:5s} {name:18s} ' + ' '.join(f'{k}={v:.2e}' if isinstance(v, float) else f'{k}={v}' for k, v in vals.items()), flush=True)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, 'w') as fh:
        json.dump(res, fh, indent=1)


if __name__ == '__main__':
    main()
