"""GPU parity tests: the CUDA path (through the C ABI) against the oracle and the golden fixtures.

Tolerances (BASELINE.json north_star): log_prob 1e-4 relative, gradients 1e-3 relative, vertices /
joints 1e-5 absolute in metres (= 1e-2 mm; the reference's own fp32 noise is ~1e-4 mm).
Gradient checks use smooth functionals and small row counts: leaky-ReLU makes the gradient
discontinuous in the pre-activations, so a single sign flip caused by fp32 rounding shows up as an
O(1e-2) difference in any implementation, the reference's own fp32-vs-fp64 included.
"""
import os

import numpy as np
import pytest
import torch

from mhentropy_b200 import MHEntHead, ManoLayer, RealNVP
from mhentropy_b200.mano_assets import synthetic_mano
from oracle import flow_oracle as fo
from oracle import loss_oracle as lo
from oracle import mano_oracle as mo

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def load(golden_dir, name):
    return dict(np.load(os.path.join(golden_dir, name)))


def T(a, dev=DEV, grad=False):
    t = torch.from_numpy(np.asarray(a)).clone().to(dev)
    return t.requires_grad_(True) if grad else t


def relerr(a, b):
    a = a.detach().cpu().double().numpy() if isinstance(a, torch.Tensor) else np.asarray(a, np.float64)
    b = b.detach().cpu().double().numpy() if isinstance(b, torch.Tensor) else np.asarray(b, np.float64)
    return np.abs(a - b).max() / (np.abs(b).max() + 1e-30)


PRECISIONS = ['fp32', 'bf16x3']


def build_flow(sd, cfg, precision='fp32'):
    flow = RealNVP(**cfg)
    flow.precision = precision
    missing = flow.load_state_dict({k: torch.as_tensor(v) for k, v in sd.items()}, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    return flow.to(DEV)


SMALL = dict(dim=45, tsfm_on=32, kemb=False, jointN=21, h_dims=[64, 64], num_steps=2)
PROD = dict(dim=45, tsfm_on=512, kemb=False, jointN=21, h_dims=[512, 512], num_steps=6)


@pytest.mark.parametrize('precision', PRECISIONS)
def test_flow_small_against_golden(golden_dir, precision):
    fx = load(golden_dir, 'flow_small.npz')
    sd = {k[2:]: v for k, v in fx.items() if k.startswith('w/')}
    flow = build_flow(sd, SMALL, precision)
    feat, z0 = T(fx['feat'], grad=True), T(fx['z0'], grad=True)
    x = flow.forward_p(z0, cond=feat)
    assert np.abs(x.detach().cpu().numpy() - fx['x']).max() < (2e-5 if precision == 'fp32' else 5e-4)
    (x * T(fx['wx'])).sum().backward()
    assert relerr(feat.grad, fx['sample_dfeat']) < 1e-3
    assert relerr(z0.grad, fx['sample_dz0']) < 1e-3
    for k, p in flow.named_parameters():
        assert relerr(p.grad, fx['gs/' + k]) < 1e-3, k
    flow.zero_grad()
    xin, feat2 = T(fx['xin'], grad=True), T(fx['feat'], grad=True)
    z, lp = flow.log_prob(xin, logvar=feat2, return_z=True)
    assert np.abs(z.detach().cpu().numpy() - fx['z']).max() < (2e-5 if precision == 'fp32' else 5e-4)
    assert relerr(lp, fx['log_prob']) < 1e-4
    (lp * T(fx['wl'])).sum().backward()
    assert relerr(xin.grad, fx['logprob_dx']) < 1e-3
    assert relerr(feat2.grad, fx['logprob_dfeat']) < 1e-3
    for k, p in flow.named_parameters():
        assert relerr(p.grad, fx['gl/' + k]) < 1e-3, k
    with torch.no_grad():
        lpx = flow.log_prob(T(fx['x']), logvar=T(fx['feat']))
        xs, lq = flow.sample_with_log_prob(T(fx['feat']), T(fx['z0']), 1)
    assert relerr(lpx, fx['log_prob_of_x']) < 1e-4
    assert relerr(lq, fx['log_prob_of_x']) < 1e-4      # fused single-pass log q == reference's second pass


@pytest.mark.parametrize('precision', PRECISIONS)
def test_flow_prod_against_golden_and_oracle(golden_dir, precision):
    fx = load(golden_dir, 'flow_prod.npz')
    sd = fo.init_state_dict(seed=int(fx['seed']))
    flow = build_flow(sd, PROD, precision)
    feat, z0 = T(fx['feat'], grad=True), T(fx['z0'], grad=True)
    x = flow.forward_p(z0, cond=feat)
    assert np.abs(x.detach().cpu().numpy() - fx['x']).max() < (5e-5 if precision == 'fp32' else 2e-3)
    (x * T(fx['wx'])).sum().backward()
    assert relerr(feat.grad, fx['sample_dfeat']) < 1e-3
    bad = []
    for k, p in flow.named_parameters():
        n = float(p.grad.double().norm())
        ref = float(fx['gsnorm/' + k])
        if abs(n - ref) > 1e-3 * ref + 1e-12:
            bad.append((k, n, ref))
    assert not bad, bad[:5]
    for k in [k for k in fx if k.startswith('gsslice/')]:
        g = dict(flow.named_parameters())[k[len('gsslice/'):]].grad
        g = g[:8, :8] if g.dim() == 2 else g
        assert relerr(g, fx[k]) < 2e-3, k
    with torch.no_grad():
        lp = flow.log_prob(T(fx['xin']), logvar=T(fx['feat']))
    assert relerr(lp, fx['log_prob']) < 1e-4


@pytest.mark.parametrize('precision', PRECISIONS)
def test_flow_hoisted_conditioning_matches_repeated_features(precision):
    sd = fo.init_state_dict(dim=45, cond_dim=32, h_dims=(64, 64), num_steps=2, seed=5)
    flow = build_flow(sd, SMALL, precision)
    g = torch.Generator().manual_seed(3)
    B, S = 3, 4
    feat = torch.randn(B, 32, generator=g).to(DEV)
    z0 = torch.randn(B * S, 45, generator=g).to(DEV)
    with torch.no_grad():
        x_hoist = flow.sample(B * S, logvar=feat, z0=z0)
        x_rep = flow.sample(B * S, logvar=feat.repeat(S, 1), z0=z0)
        x_or = fo.sample(sd, z0.cpu(), feat.cpu().repeat(S, 1))
    assert torch.allclose(x_hoist, x_rep, atol=1e-5)
    assert torch.allclose(x_hoist.cpu(), x_or, atol=2e-5 if precision == 'fp32' else 5e-4)


@pytest.mark.parametrize('precision', PRECISIONS)
def test_flow_roundtrip_and_empty(precision):
    sd = fo.init_state_dict(seed=1)
    flow = build_flow(sd, PROD, precision)
    g = torch.Generator().manual_seed(0)
    feat = torch.randn(16, 512, generator=g).to(DEV)
    z0 = torch.randn(16 * 8, 45, generator=g).to(DEV)
    with torch.no_grad():
        x = flow.forward_p(z0, cond=feat)
        z, logdet = flow.backward_p(x, cond=feat)
        x2, lq = flow.sample_with_log_prob(feat, z0, 8)
        lp = flow.log_prob(x, logvar=feat)
    assert (z - z0).abs().max() < (1e-4 if precision == 'fp32' else 2e-3)   # backward_p(forward_p(z)) == z
    assert torch.equal(x, x2)
    assert relerr(lq, lp) < 1e-4
    with torch.no_grad():
        e = flow.forward_p(torch.empty(0, 45, device=DEV), cond=feat)
    assert e.shape == (0, 45)


def test_mano_against_golden(golden_dir):
    fx = load(golden_dir, 'mano.npz')
    layer = ManoLayer(flat_hand_mean=False, ncomps=45, use_pca=True, skeidx='RHD', mano_data=synthetic_mano(0)).to(DEV)
    theta, beta = T(fx['theta'], grad=True), T(fx['beta'], grad=True)
    out = layer(beta=beta, theta=theta)
    for k in ('mesh', 'joints', 'mano_joints'):
        assert np.abs(out[k].detach().cpu().numpy() - fx[k]).max() < 1e-2, k      # mm; == 1e-5 m
        assert np.abs(out[k].detach().cpu().numpy() - fx[k]).max() < 5e-4, k      # and near the fp32 noise floor
    ((out['mesh'] * T(fx['wv'])).sum() + (out['joints'] * T(fx['wj'])).sum() + (out['mano_joints'] * T(fx['wm'])).sum()).backward()
    assert relerr(theta.grad, fx['dtheta']) < 1e-3
    assert relerr(beta.grad, fx['dbeta']) < 1e-3
    theta2, beta2 = T(fx['theta'], grad=True), T(fx['beta'], grad=True)
    out2 = layer(beta=beta2, theta=theta2, want_mesh=False)
    assert torch.allclose(out2['mano_joints'], out['mano_joints'].detach(), atol=1e-3)
    (out2['mano_joints'] * T(fx['wm'])).sum().backward()
    assert relerr(theta2.grad, fx['dtheta_jonly']) < 1e-3
    assert relerr(beta2.grad, fx['dbeta_jonly']) < 1e-3
    # inner layer: manopth order
    v, j = layer.mano_layer(T(fx['theta']), T(fx['beta']))
    c = mo.mano_constants(synthetic_mano(0))
    vo, jo = mo.mano_forward(c, torch.from_numpy(fx['theta']), torch.from_numpy(fx['beta']))
    assert (v.cpu() - vo).abs().max() < 5e-4 and (j.cpu() - jo).abs().max() < 5e-4
    assert torch.equal(layer.mano_faces.cpu(), torch.from_numpy(fx['faces']))


def test_mano_large_batch_vs_oracle():
    mano = synthetic_mano(0)
    layer = ManoLayer(flat_hand_mean=False, ncomps=45, use_pca=True, skeidx='RHD', mano_data=mano).to(DEV)
    g = torch.Generator().manual_seed(4)
    R = 1500   # exercises the 8-row tile variant and ragged tails
    theta = torch.cat([0.8 * torch.randn(R, 3, generator=g), 0.9 * torch.randn(R, 45, generator=g)], 1)
    beta = 0.03 * torch.randn(R, 10, generator=g)
    out = layer(beta=beta.to(DEV), theta=theta.to(DEV))
    ref = mo.mano_wrapper_forward(mo.mano_constants(mano, torch.float64), theta.double(), beta.double())
    for k in ('mesh', 'joints', 'mano_joints'):
        assert (out[k].cpu().double() - ref[k]).abs().max() < 1e-3, k


@pytest.mark.parametrize('name', ['mhent_small.npz', 'mhent_prod.npz'])
@pytest.mark.parametrize('fused', [True, False])
@pytest.mark.parametrize('precision', PRECISIONS)
def test_mhent_loss_against_golden(golden_dir, name, fused, precision):
    fx = load(golden_dir, name)
    small = name == 'mhent_small.npz'
    cfg = dict(h_dims=[64, 64], num_steps=2, tsfm_on=32) if small else {}
    head = MHEntHead(q_z_giv_i_cfg=cfg, mano_data=synthetic_mano(0), feat_dim=32 if small else 512)
    sd = {k[2:]: torch.as_tensor(v) for k, v in fx.items() if k.startswith('w/')} if small else fo.init_state_dict(seed=int(fx['seed']))
    head.q_z_giv_i.load_state_dict(sd)
    head.q_z_giv_i.precision = precision
    head = head.to(DEV)
    feat, z_det = T(fx['feat'], grad=True), T(fx['z_det'], grad=True)
    y = {'crop_uv': T(fx['crop_uv']), 'vis': T(fx['vis'])}
    out = head.get_loss(feat, y, z0=T(fx['z0_train']), z_det=z_det, N=10, fused=fused)
    loss = (-out['log_p']).mean()
    assert relerr(out['log_p'], fx['log_p']) < 1e-4
    assert relerr(out['h_q_z_giv_i'], fx['h_q_z_giv_i']) < 1e-4
    assert relerr(out['q_log_p_z_giv_y'], fx['q_log_p_z_giv_y']) < 1e-4
    assert relerr(out['th_norm'], fx['th_norm']) < 1e-4
    assert relerr(loss, fx['loss']) < 1e-4
    loss.backward()
    # Gradient bar, by precision (tests/_gradcheck.py): the exact-fp32 path meets 1e-3 outright; on the tensor-core path a leaky-ReLU
    # crossing (visible as ONE image's dfeat row off by more than 1e-3) may move that image's share of the gradients - every other image
    # must meet the bar, and without a crossing everything does.
    from _gradcheck import kink_aware_ok
    ok, med, n_cross = kink_aware_ok(feat.grad, T(fx['dfeat']))
    bar = 1e-3 if (precision == 'fp32' or n_cross == 0) else 5e-3
    print(f'{precision}: dfeat max-rel {relerr(feat.grad, fx["dfeat"]):.2e} (median image {med:.2e}, {n_cross} images with a crossing); bar {bar:g}')
    assert ok
    assert relerr(feat.grad, fx['dfeat']) < bar
    assert relerr(z_det.grad, fx['dz_det']) < 1e-3
    grads = dict(head.q_z_giv_i.named_parameters())
    if small:
        worst = max(relerr(p.grad, fx['g/' + k]) for k, p in grads.items())
        assert worst < bar, worst
        s = head.sample(T(fx['feat']), N=3, temp=0.8, z0=T(fx['z0_sample']), z_det=T(fx['z_det']))
        for k in ('th_bt', 'logs_t', 'xyz', 'uv', 'verts'):
            assert relerr(s[k], fx['sample/' + k]) < 1e-4, k
    else:
        num = sum(float((p.grad.double().norm() - float(fx['gnorm/' + k])) ** 2) for k, p in grads.items())
        den = sum(float(fx['gnorm/' + k]) ** 2 for k in grads)
        assert (num / den) ** 0.5 < 1e-3


@pytest.mark.parametrize('precision', PRECISIONS)
def test_training_step_vs_fp64_oracle_config1(precision):
    """Config 1 shape (B=8, S=10, production flow): CUDA vs the fp64 oracle, error no worse than a small
    multiple of the fp32 oracle's own error against fp64."""
    mano = synthetic_mano(0)
    sd = fo.init_state_dict(seed=0)
    B, S = 8, 10
    batch = lo.synthetic_batch(B, S, seed=123)

    def run_oracle(dtype):
        sdg = {k: v.to(dtype).clone().requires_grad_(k != 'mask') for k, v in sd.items()}
        c = mo.mano_constants(mano, dtype)
        feat = batch['feat'].to(dtype).clone().requires_grad_(True)
        out = lo.reverse_kld(sdg, c, feat, batch['z_det'].to(dtype), batch['z0'].to(dtype), batch['crop_uv'].to(dtype),
                             batch['vis'].to(dtype), S)
        lo.mhent_loss(out['log_p']).backward()
        return out, feat.grad, {k: v.grad for k, v in sdg.items() if k != 'mask'}

    o64, df64, g64 = run_oracle(torch.float64)
    o32, df32, g32 = run_oracle(torch.float32)
    head = MHEntHead(mano_data=mano)
    head.q_z_giv_i.load_state_dict(sd)
    head.q_z_giv_i.precision = precision
    head = head.to(DEV)
    feat = batch['feat'].detach().clone().to(DEV).requires_grad_(True)
    y = {'crop_uv': batch['crop_uv'].to(DEV), 'vis': batch['vis'].to(DEV)}
    out = head.get_loss(feat, y, z0=batch['z0'].to(DEV), z_det=batch['z_det'].to(DEV), N=S)
    (-out['log_p']).mean().backward()
    assert relerr(out['log_p'], o64['log_p']) < 1e-4
    assert relerr(out['h_q_z_giv_i'], o64['h_q_z_giv_i']) < 1e-4
    tot = lambda g: torch.cat([v.double().flatten().cpu() for v in g.values()])  # noqa: E731
    gk = {k: p.grad for k, p in head.q_z_giv_i.named_parameters()}
    ref = tot(g64)
    err_cuda = float((tot(gk) - ref).norm() / ref.norm())
    err_f32 = float((tot(g32) - ref).norm() / ref.norm())
    assert err_cuda < max(1e-3, 4 * err_f32), (err_cuda, err_f32)
    assert relerr(feat.grad, df64) < max(1e-3, 4 * relerr(df32, df64))


def test_library_is_the_path():
    """The CUDA extension is what runs: kernel launches are counted, and CPU tensors never reach it."""
    from mhentropy_b200 import _lib
    before = _lib.lib().mhe_kernel_launch_count()
    sd = fo.init_state_dict(dim=45, cond_dim=32, h_dims=(64, 64), num_steps=2, seed=5)
    flow = build_flow(sd, SMALL)
    with torch.no_grad():
        flow.forward_p(torch.randn(4, 45, device=DEV), cond=torch.randn(4, 32, device=DEV))
    assert _lib.lib().mhe_kernel_launch_count() > before
    assert _lib.lib().mhe_built_for_sm() == 100


@pytest.mark.parametrize('precision', PRECISIONS)
@pytest.mark.parametrize('fuse', [True, False])
def test_mhent_loss_with_real_det_head_against_golden(golden_dir, precision, fuse):
    """MHEntHead.get_loss with ITS det_head in the graph (feat -> det_head -> z_det -> ...), against the fixture the unmodified reference
    produced with its own det_head (network.py:376-385): log_p, dfeat (flow + det_head paths summed) and the det_head gradients."""
    from mhentropy_b200.synthetic import det_head_state_dict
    from _gradcheck import kink_aware_ok
    fx = np.load(os.path.join(golden_dir, 'mhent_dethead.npz'))
    head = MHEntHead(mano_data=synthetic_mano(0))
    head.q_z_giv_i.load_state_dict(fo.init_state_dict(seed=int(fx['seed'])))
    head.det_head.load_state_dict(det_head_state_dict(5))
    head.q_z_giv_i.precision = precision
    head.fuse_loss = fuse
    head = head.to(DEV)
    for p in head.parameters():
        p.requires_grad_(True)
    feat = T(fx['feat'], grad=True)
    out = head.get_loss(feat, {'crop_uv': T(fx['crop_uv']), 'vis': T(fx['vis'])}, z0=T(fx['z0_train']), N=10)
    loss = (-out['log_p']).mean()
    loss.backward()
    assert relerr(out['log_p'], fx['log_p']) < 1e-4
    assert relerr(loss, fx['loss']) < 1e-4
    ok, med, n_cross = kink_aware_ok(feat.grad, T(fx['dfeat']))
    bar = 1e-3 if (precision == 'fp32' or n_cross == 0) else 5e-3
    assert ok and relerr(feat.grad, fx['dfeat']) < bar
    dh = dict(head.det_head.named_parameters())
    assert relerr(dh['2.weight'].grad, fx['gdet/2.weight']) < 1e-3
    assert relerr(dh['0.weight'].grad[:16, :64], fx['gdet/0.weight']) < 1e-3
    for k in ('0.weight', '0.bias', '2.weight', '2.bias'):
        n = float(dh[k].grad.double().norm())
        assert abs(n - float(fx['gdetnorm/' + k])) < 1e-3 * float(fx['gdetnorm/' + k]), k
