"""GPU parity of hypothesis selection and the multi-hypothesis metrics through the C ABI, against the
golden outputs of reference criteria.py:MHEntLoss / torch.topk and against the oracle at the 200-sample evaluation size."""
import os

import numpy as np
import pytest
import torch

from mhentropy_b200 import MHEntLoss, hypothesis_metrics, topk_hypotheses
from oracle import metrics_oracle as meo

pytestmark = pytest.mark.gpu
DEV = 'cuda'


@pytest.mark.parametrize('tag', ['a', 'b', 'c'])
def test_metrics_against_golden(golden_dir, tag):
    fx = np.load(os.path.join(golden_dir, 'metrics.npz'))
    t = lambda k: torch.from_numpy(fx[f'{tag}/{k}']).to(DEV)  # noqa: E731
    out = {'log_p': torch.zeros(t('scale').shape[0], device=DEV), 'xyz': t('xyz'), 'uv': t('uv')}
    target = {'pose3d': t('pose3d'), 'scale': t('scale'), 'crop_uv': t('crop_uv'), 'vis': t('vis')}
    loss, losses, metrics = MHEntLoss()(out, target)
    assert float(loss) == 0.0 and sorted(metrics) == sorted(meo.METRIC_KEYS)
    for k in meo.METRIC_KEYS:
        ref = fx[f'{tag}/m/{k}']
        np.testing.assert_allclose(metrics[k].cpu().numpy(), ref, rtol=2e-5, atol=2e-6 * max(1.0, float(np.abs(ref).max())), err_msg=k)
    for name in fx.files:
        if name.startswith(f'{tag}/topk'):
            kk = int(name.split('topk')[1])
            assert np.array_equal(topk_hypotheses(t('log_q'), kk).cpu().numpy(), fx[name]), name   # index work: exact


def test_metrics_evaluation_size_against_oracle():
    """N = 200 hypotheses (CrossModalHand.py:357-361) x B = 64 images, fp64 oracle."""
    g = torch.Generator().manual_seed(5)
    N, B = 200, 64
    pose3d = torch.randn(B, 63, generator=g)
    xyz = pose3d[None] + 0.3 * torch.randn(N, B, 63, generator=g)
    crop_uv = torch.rand(B, 42, generator=g) * 2 - 1
    uv = (crop_uv[None] + 1) / 2 * 256 + 8. * torch.randn(N, B, 42, generator=g)
    scale = 0.05 + 0.1 * torch.rand(B, generator=g)
    vis = (torch.rand(B, 21, generator=g) < 0.7).float()
    ref = meo.hypothesis_metrics(*[a.double() for a in (xyz, uv, pose3d, scale, crop_uv, vis)])
    out = hypothesis_metrics(*[a.to(DEV) for a in (xyz, uv, pose3d, scale, crop_uv, vis)])
    for k in meo.METRIC_KEYS:
        np.testing.assert_allclose(out[k].cpu().double().numpy(), ref[k].numpy(), rtol=1e-4, err_msg=k)
    log_q = torch.randn(N, B, generator=g)
    log_q[3] = log_q[7]                                   # ties keep the lower index first, like torch.topk on CPU
    idx = topk_hypotheses(log_q.to(DEV), 20).cpu()
    vals = torch.gather(log_q, 0, idx)
    assert torch.equal(vals, torch.topk(log_q, 20, dim=0)[0])         # selected values: exact and sorted
    assert all(len(set(idx[:, b].tolist())) == 20 for b in range(B))


def test_sample_topk_selection_uses_kernel():
    """MHEntHead.sample with N_quant < N returns the N_quant most likely hypotheses (network.py:866-871)."""
    from mhentropy_b200 import MHEntHead
    from mhentropy_b200.mano_assets import synthetic_mano
    torch.manual_seed(0)
    head = MHEntHead(mano_data=synthetic_mano(0)).to(DEV)
    B, N, Q = 4, 12, 5
    feat = torch.randn(B, 512, device=DEV)
    z_det = torch.cat([0.5 * torch.randn(B, 3), 0.02 * torch.randn(B, 10), torch.full((B, 1), -1.2), 0.1 * torch.randn(B, 2)], 1).to(DEV)
    z0 = torch.randn(N * B, 45, device=DEV) * 0.8
    full = head.sample(feat, N=N, temp=0.8, mods=['xyz'], z0=z0, z_det=z_det)
    sel = head.sample(feat, N=[N, Q], temp=0.8, mods=['xyz'], z0=z0, z_det=z_det)
    log_q = head._reverse_log_q(torch.cat([full['th_bt'], full['logs_t']], -1).flatten(0, 1), feat).reshape(N, B)
    idx = torch.topk(log_q, Q, dim=0)[1]
    want = torch.gather(full['th_bt'], 0, idx[..., None].repeat(1, 1, 58))
    assert sel['th_bt'].shape == (Q, B, 58)
    assert torch.equal(sel['th_bt'], want)
