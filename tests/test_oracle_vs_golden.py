"""Pin the oracle (oracle/*.py) to the reference: every golden fixture under tests/golden/ was
produced by the unmodified reference (tests/golden/make_golden.py); where /root/reference is
mounted the live reference is checked as well."""
import os

import numpy as np
import pytest
import torch

from mhentropy_b200.mano_assets import synthetic_mano
from oracle import flow_oracle as fo
from oracle import loss_oracle as lo
from oracle import mano_oracle as mo
from oracle import ref_shim


def load(golden_dir, name):
    return dict(np.load(os.path.join(golden_dir, name)))


def T(a, grad=False):
    t = torch.from_numpy(np.asarray(a)).clone()
    return t.requires_grad_(True) if grad else t


def state_dict_from(fx, prefix='w/'):
    return {k[len(prefix):]: T(v) for k, v in fx.items() if k.startswith(prefix)}


def _np(a):
    if isinstance(a, torch.Tensor):
        a = a.detach().cpu().numpy()
    return np.asarray(a, dtype=np.float64)


def l2(a):
    return float(np.sqrt((_np(a) ** 2).sum()))


def relerr(a, b):
    a, b = _np(a), _np(b)
    return np.abs(a - b).max() / (np.abs(b).max() + 1e-30)


def test_flow_small_forward_and_grads(golden_dir):
    fx = load(golden_dir, 'flow_small.npz')
    sd = state_dict_from(fx)
    sdg = {k: (v.requires_grad_(True) if k != 'mask' else v) for k, v in sd.items()}
    feat, z0 = T(fx['feat'], True), T(fx['z0'], True)
    x = fo.sample(sdg, z0, feat)
    assert np.abs(x.detach().numpy() - fx['x']).max() <= 1e-6
    (x * T(fx['wx'])).sum().backward()
    assert relerr(feat.grad, fx['sample_dfeat']) < 1e-5
    assert relerr(z0.grad, fx['sample_dz0']) < 1e-5
    for k, v in sdg.items():
        if k != 'mask':
            assert relerr(v.grad, fx['gs/' + k]) < 1e-4, k
            v.grad = None
    xin, feat2 = T(fx['xin'], True), T(fx['feat'], True)
    z, lp = fo.log_prob(sdg, xin, feat2, return_z=True)
    assert np.abs(z.detach().numpy() - fx['z']).max() <= 1e-6
    assert relerr(lp.detach(), fx['log_prob']) < 1e-6
    (lp * T(fx['wl'])).sum().backward()
    assert relerr(xin.grad, fx['logprob_dx']) < 1e-5
    assert relerr(feat2.grad, fx['logprob_dfeat']) < 1e-5
    for k, v in sdg.items():
        if k != 'mask':
            assert relerr(v.grad, fx['gl/' + k]) < 1e-4, k
    # entropy identity: log q(x) of own samples == log N(z0) - sum s  
    with torch.no_grad():
        xs, logdet = fo.forward_p(sd, T(fx['z0']), T(fx['feat']), return_logdet=True)
        ident = fo.std_normal_log_prob(T(fx['z0'])) - logdet
    assert relerr(ident, fx['log_prob_of_x']) < 1e-5


def test_flow_prod_weights_from_seed(golden_dir):
    fx = load(golden_dir, 'flow_prod.npz')
    sd = fo.init_state_dict(seed=int(fx['seed']))
    n_params = sum(v.numel() for k, v in sd.items() if k != 'mask')
    assert n_params == 20_030_520
    sdg = {k: (v.requires_grad_(True) if k != 'mask' else v) for k, v in sd.items()}
    feat, z0 = T(fx['feat'], True), T(fx['z0'], True)
    x = fo.sample(sdg, z0, feat)
    assert np.abs(x.detach().numpy() - fx['x']).max() <= 2e-6
    (x * T(fx['wx'])).sum().backward()
    assert relerr(feat.grad, fx['sample_dfeat']) < 1e-5
    for k, v in sdg.items():
        if k != 'mask':
            n = l2(v.grad)
            assert abs(n - float(fx['gsnorm/' + k])) <= 1e-4 * float(fx['gsnorm/' + k]) + 1e-12, k
    for k in [k for k in fx if k.startswith('gsslice/')]:
        name = k[len('gsslice/'):]
        g = sdg[name].grad
        g = g[:8, :8] if g.dim() == 2 else g
        assert relerr(g, fx[k]) < 1e-4, k
    lp = fo.log_prob(sd, T(fx['xin']), T(fx['feat']))
    assert relerr(lp, fx['log_prob']) < 1e-6


def test_mano_forward_and_grads(golden_dir):
    fx = load(golden_dir, 'mano.npz')
    c = mo.mano_constants(synthetic_mano(0))
    theta, beta = T(fx['theta'], True), T(fx['beta'], True)
    out = mo.mano_wrapper_forward(c, theta, beta)
    # mm; reference fp32 noise floor is ~1e-4 mm at |v| ~ 200 mm
    assert np.abs(out['mesh'].detach().numpy() - fx['mesh']).max() < 2e-4
    assert np.abs(out['joints'].detach().numpy() - fx['joints']).max() < 2e-4
    assert np.abs(out['mano_joints'].detach().numpy() - fx['mano_joints']).max() < 2e-4
    ((out['mesh'] * T(fx['wv'])).sum() + (out['joints'] * T(fx['wj'])).sum()
     + (out['mano_joints'] * T(fx['wm'])).sum()).backward()
    assert relerr(theta.grad, fx['dtheta']) < 1e-4
    assert relerr(beta.grad, fx['dbeta']) < 1e-4
    # fp64 oracle agrees with the fp32 reference to fp32 round-off
    c64 = mo.mano_constants(synthetic_mano(0), torch.float64)
    out64 = mo.mano_wrapper_forward(c64, T(fx['theta']).double(), T(fx['beta']).double())
    assert np.abs(out64['mesh'].numpy() - fx['mesh']).max() < 2e-4
    # tips of mano_joints are exactly mesh vertices
    mj = out['mano_joints'].detach()[:, mo.FREIHAND2RHD]  # FREIHAND2RHD is an involution
    for slot, vid in zip((4, 8, 12, 16, 20), mo.TIP_VERTS):
        j = mo.JOINT_REORDER.index(16 + (slot // 4 - 1))
        assert torch.equal(mj[:, j], out['mesh'].detach()[:, vid])


@pytest.mark.parametrize('name,tol_w', [('mhent_small.npz', 1e-4), ('mhent_prod.npz', 1e-4)])
def test_mhent_loss(golden_dir, name, tol_w):
    fx = load(golden_dir, name)
    if name == 'mhent_small.npz':
        sd = state_dict_from(fx)
    else:
        sd = fo.init_state_dict(seed=int(fx['seed']))
    c = mo.mano_constants(synthetic_mano(0))
    sdg = {k: (v.requires_grad_(True) if k != 'mask' else v) for k, v in sd.items()}
    feat, z_det = T(fx['feat'], True), T(fx['z_det'], True)
    N = fx['z0_train'].shape[0] // fx['feat'].shape[0]
    out = lo.reverse_kld(sdg, c, feat, z_det, T(fx['z0_train']), T(fx['crop_uv']), T(fx['vis']), N)
    loss = lo.mhent_loss(out['log_p'])
    assert relerr(out['log_p'].detach(), fx['log_p']) < 1e-6
    assert relerr(out['h_q_z_giv_i'].detach(), fx['h_q_z_giv_i']) < 1e-6
    assert relerr(out['th_norm'].detach(), fx['th_norm']) < 1e-6
    assert relerr(loss.detach(), fx['loss']) < 1e-6
    loss.backward()
    assert relerr(feat.grad, fx['dfeat']) < 1e-4
    assert relerr(z_det.grad, fx['dz_det']) < 1e-4
    if name == 'mhent_small.npz':
        for k, v in sdg.items():
            if k != 'mask':
                assert relerr(v.grad, fx['g/' + k]) < tol_w, k
        Ns = fx['z0_sample'].shape[0] // fx['feat'].shape[0]
        with torch.no_grad():
            s = lo.mhent_sample(sd, c, T(fx['feat']), T(fx['z_det']), T(fx['z0_sample']), Ns)
        for k in ('th_bt', 'logs_t', 'xyz', 'uv'):
            assert relerr(s[k], fx['sample/' + k]) < 1e-5, k
        assert relerr(s['verts'], fx['sample/verts']) < 1e-5
    else:
        for k, v in sdg.items():
            if k != 'mask':
                ref = float(fx['gnorm/' + k])
                n = l2(v.grad)
                assert abs(n - ref) <= tol_w * ref + 1e-12, k


@pytest.mark.skipif(not ref_shim.reference_available(), reason='/root/reference not mounted')
def test_oracle_vs_live_reference_flow():
    mano = synthetic_mano(0)
    flows = ref_shim.import_flows(mano)
    cfg = dict(dim=45, tsfm_on=16, kemb=False, jointN=21, h_dims=[32, 48], num_steps=3)
    with ref_shim.cpu_mode():
        torch.manual_seed(21)
        flow = flows.RealNVP(**cfg)
        sd = {k: v.detach().clone() for k, v in flow.state_dict().items()}
        feat, z0 = torch.randn(7, 16), torch.randn(7, 45)
        x_ref = flow.forward_p(z0, cond=feat)
        lp_ref = flow.log_prob(x_ref, logvar=feat)
    sd2 = fo.init_state_dict(dim=45, cond_dim=16, h_dims=(32, 48), num_steps=3, seed=21)
    assert all(torch.equal(sd[k], sd2[k]) for k in sd)
    assert torch.allclose(fo.sample(sd, z0, feat), x_ref, atol=1e-6)
    assert torch.allclose(fo.log_prob(sd, x_ref, feat), lp_ref, rtol=1e-6)


@pytest.mark.parametrize('tag', ['a', 'b', 'c'])
def test_metrics_oracle_against_golden(golden_dir, tag):
    """oracle/metrics_oracle.py against reference criteria.py:MHEntLoss outputs (tests/golden/metrics.npz): occluded /
    all-visible / none-visible images (a), N = 1 (b), an empty joint group for the whole batch (c)."""
    from oracle import metrics_oracle as meo
    fx = np.load(os.path.join(golden_dir, 'metrics.npz'))
    t = lambda k: torch.from_numpy(fx[f'{tag}/{k}'])  # noqa: E731
    out = meo.hypothesis_metrics(t('xyz'), t('uv'), t('pose3d'), t('scale'), t('crop_uv'), t('vis'))
    assert sorted(out) == sorted(meo.METRIC_KEYS)
    for k in meo.METRIC_KEYS:
        ref = fx[f'{tag}/m/{k}']
        np.testing.assert_allclose(out[k].numpy(), ref, rtol=1e-5, atol=1e-6 * max(1.0, float(np.abs(ref).max())))
    for name in fx.files:
        if name.startswith(f'{tag}/topk'):
            kk = int(name.split('topk')[1])
            assert np.array_equal(meo.topk_hypotheses(t('log_q'), kk).numpy(), fx[name])


def test_oracle_with_real_det_head_against_golden(golden_dir):
    """The fixture ran the reference's OWN det_head (network.py:376-385) inside MHEnt.get_loss; the oracle composes the same head
    (plain torch Linear - ReLU - Linear on the seeded weights) with reverse_kld and must reproduce log_p and the gradients."""
    from mhentropy_b200.synthetic import det_head_state_dict
    fx = load(golden_dir, 'mhent_dethead.npz')
    sd = fo.init_state_dict(seed=int(fx['seed']))
    sdg = {k: v.clone().requires_grad_(k != 'mask') for k, v in sd.items()}
    dh = {k: v.clone().requires_grad_(True) for k, v in det_head_state_dict(5).items()}
    feat = T(fx['feat'], True)
    z_det = torch.relu(feat @ dh['0.weight'].T + dh['0.bias']) @ dh['2.weight'].T + dh['2.bias']
    assert relerr(z_det, fx['z_det']) < 1e-5
    out = lo.reverse_kld(sdg, mo.mano_constants(synthetic_mano(0)), feat, z_det, T(fx['z0_train']), T(fx['crop_uv']), T(fx['vis']), 10)
    lo.mhent_loss(out['log_p']).backward()
    assert relerr(out['log_p'], fx['log_p']) < 1e-5
    assert relerr(feat.grad, fx['dfeat']) < 1e-4
    assert relerr(dh['2.weight'].grad, fx['gdet/2.weight']) < 1e-4
    assert relerr(dh['0.weight'].grad[:16, :64], fx['gdet/0.weight']) < 1e-4
    for k in ('0.weight', '0.bias', '2.weight', '2.bias'):
        assert abs(l2(dh[k].grad) - float(fx['gdetnorm/' + k])) < 1e-4 * float(fx['gdetnorm/' + k]), k
