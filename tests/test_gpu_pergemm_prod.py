"""The per-GEMM tensor-core path (csrc/flow_tc.cu) at PRODUCTION width against the fp64 oracle.

It carries every row count above the cluster-fused kernels' limit (MHE_FUSED_MAX_ROWS = 4096): BASELINE configs 3 (25,600 rows),
4 (16,384) and 5 (32,768 per GPU).  Forward, inverse and the whole backward are checked at H = 512 with ragged last tiles, once with
the conditioning hoisted per image (<= 128 images: the fp32-streaming conditioning GEMM) and once config-4 style with one feature row
per pose (> 128 "images": the conditioning GEMMs on half / bfloat16 planes).  Reference: hand/flows.py:210-227, 271-331.
"""
import numpy as np
import pytest
import torch

from mhentropy_b200 import RealNVP, _lib
from oracle import flow_oracle as fo
from _gradcheck import flat_error, fro

pytestmark = pytest.mark.gpu
DEV = 'cuda'
PROD = dict(dim=45, tsfm_on=512, kemb=False, jointN=21, h_dims=[512, 512], num_steps=6)


def relmax(a, b):
    a, b = a.detach().cpu().double().numpy(), b.detach().cpu().double().numpy()
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30))


@pytest.fixture(scope='module')
def flow_and_sd():
    sd = fo.init_state_dict(seed=0)
    flow = RealNVP(**PROD)
    flow.precision = 'bf16x3'
    flow.load_state_dict(sd, strict=True)
    flow = flow.to(DEV)
    for p in flow.parameters():
        p.requires_grad_(True)
    return flow, sd


def _oracle(sd, feat_rows, inp, wx, wl, direction, dtype):
    """fp64 / fp32 autograd of the oracle: direction 0 = sample + log q of the samples, 1 = log_prob of given poses."""
    sdg = {k: v.to(dtype).clone().requires_grad_(k != 'mask') for k, v in sd.items()}
    f = feat_rows.to(dtype).clone().requires_grad_(True)
    z = inp.to(dtype).clone().requires_grad_(True)
    if direction == 0:
        S = inp.shape[0] // feat_rows.shape[0]
        out, ld = fo.forward_p(sdg, z, f.repeat(S, 1), return_logdet=True)
        lp = fo.std_normal_log_prob(z) - ld
    else:
        out, lp = fo.log_prob(sdg, z, f, return_z=True)
    ((out * wx.to(dtype)).sum() + (lp * wl.to(dtype)).sum()).backward()
    return {'out': out.detach(), 'lp': lp.detach(), 'din': z.grad, 'dfeat': f.grad, 'params': {k: v.grad for k, v in sdg.items() if k != 'mask'}}


def _check(got, ref, o32, what):
    e_out = float((got['out'].detach().cpu().double() - ref['out']).abs().max())
    e_lp = relmax(got['lp'], ref['lp'])
    e_din, e_df = fro(got['din'], ref['din']), fro(got['dfeat'], ref['dfeat'])
    e_flat, n_over, worst_name, worst = flat_error(got['params'], ref['params'])
    f_din, f_df, f_flat = fro(o32['din'], ref['din']), fro(o32['dfeat'], ref['dfeat']), flat_error(o32['params'], ref['params'])[0]
    print(f'{what}: |d out| {e_out:.2e}  log-prob rel {e_lp:.2e}  d-input {e_din:.2e} (fp32 ref {f_din:.2e})  dfeat {e_df:.2e} ({f_df:.2e})  '
          f'flat gradient {e_flat:.2e} ({f_flat:.2e}); {n_over} of 240 tensors above 1e-3, worst {worst_name} {worst:.2e}')
    assert e_out < 5e-4 and e_lp < 1e-4                                  # north star: log_prob within 1e-4 relative
    assert e_din < max(1e-3, 2 * f_din) and e_df < max(1e-3, 2 * f_df)   # gradients within 1e-3 (or twice the reference's fp32 floor)
    assert e_flat < max(1e-3, 2 * f_flat)


def test_pergemm_sample_forward_backward_hoisted_conditioning(flow_and_sd):
    """Sampling direction with log q, 66 images x 65 hypotheses = 4,290 rows (ragged against 64- and 128-row tiles)."""
    flow, sd = flow_and_sd
    B, S = 66, 65
    R = B * S
    g = torch.Generator().manual_seed(11)
    feat, z0 = torch.randn(B, 512, generator=g), torch.randn(R, 45, generator=g)
    wx, wl = torch.randn(R, 45, generator=g), torch.randn(R, generator=g)
    ref = _oracle(sd, feat, z0, wx, wl, 0, torch.float64)
    o32 = _oracle(sd, feat, z0, wx, wl, 0, torch.float32)
    flow.zero_grad(set_to_none=True)
    fc, zc = feat.to(DEV).requires_grad_(True), z0.to(DEV).requires_grad_(True)
    n0 = _lib.lib().mhe_kernel_launch_count()
    x, logq = flow.sample_with_log_prob(fc, zc, S)
    ((x * wx.to(DEV)).sum() + (logq * wl.to(DEV)).sum()).backward()
    torch.cuda.synchronize()
    assert _lib.lib().mhe_kernel_launch_count() - n0 > 100, 'this row count must take the per-GEMM path'
    _check({'out': x, 'lp': logq, 'din': zc.grad, 'dfeat': fc.grad, 'params': {k: p.grad for k, p in flow.named_parameters()}}, ref, o32,
           'per-GEMM sample 66 x 65')


def test_pergemm_log_prob_forward_backward_one_feature_row_per_pose(flow_and_sd):
    """Config-4 style: log_prob of 4,290 poses, each with its own feature row (the conditioning GEMMs run on planes: > 128 images)."""
    flow, sd = flow_and_sd
    R = 4290
    assert _lib.lib().mhe_flow_cond_fwd_uses_planes(flow._shape, R) == 1
    g = torch.Generator().manual_seed(12)
    feat, x = torch.randn(R, 512, generator=g), 0.5 * torch.randn(R, 45, generator=g)
    wz, wl = torch.randn(R, 45, generator=g), torch.randn(R, generator=g)
    ref = _oracle(sd, feat, x, wz, wl, 1, torch.float64)
    o32 = _oracle(sd, feat, x, wz, wl, 1, torch.float32)
    flow.zero_grad(set_to_none=True)
    fc, xc = feat.to(DEV).requires_grad_(True), x.to(DEV).requires_grad_(True)
    z, lp = flow.log_prob(xc, logvar=fc, return_z=True)
    ((z * wz.to(DEV)).sum() + (lp * wl.to(DEV)).sum()).backward()
    torch.cuda.synchronize()
    _check({'out': z, 'lp': lp, 'din': xc.grad, 'dfeat': fc.grad, 'params': {k: p.grad for k, p in flow.named_parameters()}}, ref, o32,
           'per-GEMM log_prob 4290 x 1')


def test_pergemm_log_prob_16384_rows(flow_and_sd):
    """BASELINE configs[3] at full size: 16,384 poses.  Rows are independent given the weights, so the fp64 oracle scores three row
    blocks (first, middle, last tile) and a sample-then-score round trip covers every row."""
    flow, sd = flow_and_sd
    R = 16384
    g = torch.Generator().manual_seed(13)
    feat, x = torch.randn(R, 512, generator=g), 0.5 * torch.randn(R, 45, generator=g)
    sd64 = fo.cast_state_dict(sd, torch.float64)
    with torch.no_grad():
        fc = feat.to(DEV)
        z, lp = flow.log_prob(x.to(DEV), logvar=fc, return_z=True)
        for lo_ in (0, 8000, R - 256):
            rows = slice(lo_, lo_ + 256)
            z_ref, lp_ref = fo.log_prob(sd64, x[rows].double(), feat[rows].double(), return_z=True)
            assert float((z[rows].cpu().double() - z_ref).abs().max()) < 5e-4
            assert relmax(lp[rows], lp_ref) < 1e-4
        # size-independent property over all rows: flow(flow^-1(x)) == x and the two log-determinants cancel
        xr, ld_f = flow._pass(z, fc, 0)
        _, ld_b = flow._pass(x.to(DEV), fc, 1)
        assert float((xr.cpu() - x).abs().max()) < 1e-3
        assert float((ld_f + ld_b).abs().max()) < 1e-3
    print(f'per-GEMM log_prob 16384 rows: round trip {float((xr.cpu() - x).abs().max()):.2e}')


def test_rowcond_pass_matches_materialised_conditioning_and_oracle(flow_and_sd, monkeypatch):
    """Per-row conditioning without a backward runs `mhe_flow_pass_fwd_rowcond` (the projections contracted inside the coupling GEMMs as
    a second operand pair).  At a ragged 4,290 rows, both directions: equal to the pass on materialised projections within fp32 rounding
    of a different summation order, and to the fp64 oracle within the north-star tolerances."""
    flow, sd = flow_and_sd
    R = 4290
    L = _lib.lib()
    assert L.mhe_flow_rowcond_supported(flow._shape, R) == 1 and L.mhe_flow_rowcond_supported(flow._shape, 640) == 0
    assert L.mhe_flow_rowcond_supported(flow._shape, 2048) == 1      # (MHE_ROWCOND_MIN_ROWS: 1536 by default)
    g = torch.Generator().manual_seed(21)
    feat, x = torch.randn(R, 512, generator=g), 0.5 * torch.randn(R, 45, generator=g)
    sd64 = fo.cast_state_dict(sd, torch.float64)
    with torch.no_grad():
        fc, xc = feat.to(DEV), x.to(DEV)
        assert flow._rowcond_applies(xc, fc)
        n0 = L.mhe_kernel_launch_count()
        got = {d: flow._pass(xc, fc, d) for d in (0, 1)}
        n_row = L.mhe_kernel_launch_count() - n0
        monkeypatch.setenv('MHE_FLOW_ROWCOND', '0')
        assert not flow._rowcond_applies(xc, fc)
        n0 = L.mhe_kernel_launch_count()
        mat = {d: flow._pass(xc, fc, d) for d in (0, 1)}
        n_mat = L.mhe_kernel_launch_count() - n0
        monkeypatch.delenv('MHE_FLOW_ROWCOND')
        torch.cuda.synchronize()
        assert n_row > 0 and n_mat > 0
        for d in (0, 1):
            assert float((got[d][0] - mat[d][0]).abs().max()) < 2e-4 and float((got[d][1] - mat[d][1]).abs().max()) < 2e-3
        z_ref, lp_ref = fo.log_prob(sd64, x.double(), feat.double(), return_z=True)
        z, ld = got[1]
        lp = -0.5 * (z.double() ** 2).sum(1) - 0.5 * 45 * np.log(2 * np.pi) + ld.double()
        assert float((z.cpu().double() - z_ref).abs().max()) < 5e-4 and relmax(lp, lp_ref) < 1e-4
        x_ref = fo.forward_p(sd64, x.double(), feat.double())
        assert float((got[0][0].cpu().double() - x_ref).abs().max()) < 5e-4
    # a tensor that needs gradients keeps the differentiable path
    assert not flow._rowcond_applies(xc.clone().requires_grad_(True), fc)


def test_wide_mask_takes_the_per_gemm_path_and_matches_the_oracle():
    """A 30/15 coupling split is wider than the cluster-fused kernels' 24-dim exchange: it must run (on the per-GEMM path) and agree with
    the oracle evaluated with the same mask - not silently drop dims."""
    sd = fo.init_state_dict(seed=0)
    m = torch.tensor([[0.] * 15 + [1.] * 30, [1.] * 30 + [0.] * 15] * 6)
    sd = dict(sd, mask=m)
    flow = RealNVP(mask=m.clone(), **PROD)
    flow.precision = 'bf16x3'
    flow.load_state_dict(sd)
    flow = flow.to(DEV)
    assert flow._shape.max_split == 30
    B, S = 8, 5
    g = torch.Generator().manual_seed(14)
    feat, z0 = torch.randn(B, 512, generator=g), torch.randn(B * S, 45, generator=g)
    sd64 = fo.cast_state_dict(sd, torch.float64)
    with torch.no_grad():
        x_ref, ld_ref = fo.forward_p(sd64, z0.double(), feat.repeat(S, 1).double(), return_logdet=True)
        n0 = _lib.lib().mhe_kernel_launch_count()
        x, logq = flow.sample_with_log_prob(feat.to(DEV), z0.to(DEV), S)
        torch.cuda.synchronize()
        assert _lib.lib().mhe_kernel_launch_count() - n0 > 40, 'wide splits must not take the cluster-fused kernels'
        lq_ref = fo.std_normal_log_prob(z0.double()) - ld_ref
        assert float((x.cpu().double() - x_ref).abs().max()) < 5e-4
        assert relmax(logq, lq_ref) < 1e-4
