"""Generate the golden fixtures in this directory by running the UNMODIFIED reference.

Run in the build container only (needs ``/root/reference``):

    python tests/golden/make_golden.py

The reference holds no golden vectors of its own for this path, so these are
outputs of the reference's code itself — ``RealNVP`` (``hand/flows.py``), ``ManoLayer``
(``hand/ManoLayer.py`` / ``hand/manopth``) and ``MHEnt.get_loss`` / ``MHEnt.sample``
(``hand/network.py``) — on seeded synthetic inputs and synthetic MANO-shaped constants
(``mhentropy_b200.mano_assets.synthetic_mano(0)``), imported through ``oracle/ref_shim.py``.

Fixtures
  flow_small.npz   small flow (h=64, cond=32, 4 layers) incl. weights: sample / log_prob / grads
  flow_prod.npz    production flow (h=512, cond=512, 12 layers), weights re-derived from seed 0
  mano.npz         ManoLayer forward + gradients of a fixed linear functional
  mhent_small.npz  MHEnt.get_loss fwd+bwd and MHEnt.sample with the small flow
  mhent_prod.npz   MHEnt.get_loss fwd+bwd with the production flow (B=4, N=10)
  mhent_dethead.npz  MHEnt.get_loss fwd+bwd with the production flow AND the reference's real det_head (network.py:376-385; the other
                   MHEnt fixtures replace it by a fixed z_det): weights set from mhentropy_b200.synthetic.det_head_state_dict(5)
  metrics.npz      MHEntLoss.forward metrics (criteria.py:47-173) and torch.topk selection (network.py:866-871) on
                   seeded (N, B, .) outputs: N = 7 with occluded / all-visible / none-visible images, and N = 1
                   (``python tests/golden/make_golden.py metrics`` regenerates only this one)
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from mhentropy_b200.mano_assets import synthetic_mano  # noqa: E402
from oracle import ref_shim  # noqa: E402
from oracle.loss_oracle import synthetic_batch  # noqa: E402

SMALL = dict(dim=45, tsfm_on=32, kemb=False, jointN=21, h_dims=[64, 64], num_steps=2)
PROD = dict(dim=45, tsfm_on=512, kemb=False, jointN=21, h_dims=[512, 512], num_steps=6)


def npy(t):
    return t.detach().cpu().numpy()


def flow_fixture(flows, cfg, seed, R, with_weights):
    with ref_shim.cpu_mode():
        torch.manual_seed(seed)
        flow = flows.RealNVP(**cfg)
        g = torch.Generator().manual_seed(seed + 100)
        feat = torch.randn(R, cfg['tsfm_on'], generator=g).requires_grad_(True)
        z0 = torch.randn(R, cfg['dim'], generator=g).requires_grad_(True)
        x = flow.forward_p(z0, cond=flow.make_cond(feat))
        # gradient of a fixed linear functional of the sample
        wx = torch.randn(R, cfg['dim'], generator=g)
        (x * wx).sum().backward()
        out = {'feat': npy(feat), 'z0': npy(z0), 'x': npy(x), 'wx': npy(wx),
               'sample_dfeat': npy(feat.grad), 'sample_dz0': npy(z0.grad)}
        sample_grads = {k: npy(p.grad) for k, p in flow.named_parameters()}
        flow.zero_grad()
        xin = (0.5 * torch.randn(R, cfg['dim'], generator=g)).requires_grad_(True)
        feat2 = feat.detach().clone().requires_grad_(True)
        z, lp = flow.log_prob(xin, logvar=feat2, return_z=True)
        wl = torch.randn(R, generator=g)
        (lp * wl).sum().backward()
        out.update({'xin': npy(xin), 'z': npy(z), 'log_prob': npy(lp), 'wl': npy(wl),
                    'logprob_dx': npy(xin.grad), 'logprob_dfeat': npy(feat2.grad)})
        logprob_grads = {k: npy(p.grad) for k, p in flow.named_parameters()}
        # log_prob of own samples (entropy identity)
        out['log_prob_of_x'] = npy(flow.log_prob(x.detach(), logvar=feat.detach()))
    if with_weights:
        for k, v in flow.state_dict().items():
            out['w/' + k] = npy(v)
        for k, v in sample_grads.items():
            out['gs/' + k] = v
        for k, v in logprob_grads.items():
            out['gl/' + k] = v
    else:
        # production size: weights come from the seed; keep per-tensor gradient norms and a slice
        for tag, grads in (('gs', sample_grads), ('gl', logprob_grads)):
            for k, v in grads.items():
                out[f'{tag}norm/' + k] = np.array(np.sqrt((v.astype(np.float64) ** 2).sum()))
            for k in ('s.0.l.1.weight', 't.11.l.1.weight', 's.5.c.0.weight', 't.7.l.2.weight', 's.3.l.0.weight'):
                out[f'{tag}slice/' + k] = grads[k][:8, :8].copy()
            for k in ('s.2.l.2.bias', 't.9.c.1.bias'):
                out[f'{tag}slice/' + k] = grads[k].copy()
    out['seed'] = np.array(seed)
    return out


def mano_fixture(ref_manolayer, R=6):
    with ref_shim.cpu_mode():
        layer = ref_manolayer.ManoLayer(flat_hand_mean=False, ncomps=45, use_pca=True, skeidx='RHD',
                                        output_size=256, mask_sz=64)
        g = torch.Generator().manual_seed(7)
        theta = torch.cat([0.8 * torch.randn(R, 3, generator=g), 0.9 * torch.randn(R, 45, generator=g)], dim=1)
        theta[0, :3] = 0.0   # exercises the ||v + 1e-8|| small-angle path of batch_rodrigues
        theta = theta.requires_grad_(True)
        beta = (0.03 * torch.randn(R, 10, generator=g)).requires_grad_(True)
        out = layer(beta=beta, theta=theta)
        wv = torch.randn(R, 778, 3, generator=g) * 1e-3
        wj = torch.randn(R, 21, 3, generator=g) * 1e-2
        wm = torch.randn(R, 21, 3, generator=g) * 1e-2
        ((out['mesh'] * wv).sum() + (out['joints'] * wj).sum() + (out['mano_joints'] * wm).sum()).backward()
        fx = {'theta': npy(theta), 'beta': npy(beta), 'mesh': npy(out['mesh']), 'joints': npy(out['joints']),
              'mano_joints': npy(out['mano_joints']), 'wv': npy(wv), 'wj': npy(wj), 'wm': npy(wm),
              'dtheta': npy(theta.grad), 'dbeta': npy(beta.grad), 'faces': npy(layer.mano_faces)}
        # joints-only functional (the training path: verts carry no gradient)
        theta2 = theta.detach().clone().requires_grad_(True)
        beta2 = beta.detach().clone().requires_grad_(True)
        out2 = layer(beta=beta2, theta=theta2)
        (out2['mano_joints'] * wm).sum().backward()
        fx['dtheta_jonly'] = npy(theta2.grad)
        fx['dbeta_jonly'] = npy(beta2.grad)
    return fx


class _ZDet(torch.nn.Module):
    """Stands in for ``det_head`` (feature-producer side): returns a fixed (B,16) parameter."""

    def __init__(self, z_det):
        super().__init__()
        self.z = torch.nn.Parameter(z_det.clone())

    def forward(self, feat):
        return self.z


def mhent_fixture(mano, cfg, B, N, seed, with_weight_grads, with_sample):
    model = ref_shim.build_mhent(mano, seed=seed, flow_cfg=cfg)
    batch = synthetic_batch(B, N, seed=seed + 1, cond_dim=cfg['tsfm_on'])
    model.det_head = _ZDet(batch['z_det'])
    y = {'crop_uv': batch['crop_uv'], 'vis': batch['vis'], 'st': torch.zeros(B, 3), 'image': np.zeros(1)}
    fx = {k: npy(v) for k, v in batch.items()}
    with ref_shim.cpu_mode():
        feat = batch['feat'].clone().requires_grad_(True)
        torch.manual_seed(seed + 2)
        z0 = torch.randn(N * B, 45)          # == prior.sample((N*B,)) * 1.0 under this seed
        torch.manual_seed(seed + 2)
        out = model.get_loss(feat, y, mods=['uv'])
        loss = (-out['log_p']).mean()         # MHEntLoss, criteria.py:55,173
        loss.backward()
        fx.update({'z0_train': npy(z0), 'log_p': npy(out['log_p']), 'h_q_z_giv_i': npy(out['h_q_z_giv_i']),
                   'q_log_p_z_giv_y': npy(out['q_log_p_z_giv_y']), 'th_norm': npy(out['th_norm']),
                   'bt_norm': npy(out['bt_norm']), 'loss': npy(loss), 'dfeat': npy(feat.grad),
                   'dz_det': npy(model.det_head.z.grad)})
        grads = {k: npy(p.grad) for k, p in model.q_z_giv_i.named_parameters()}
        if with_weight_grads:
            for k, v in model.q_z_giv_i.state_dict().items():
                fx['w/' + k] = npy(v)
            for k, v in grads.items():
                fx['g/' + k] = v
        else:
            for k, v in grads.items():
                fx['gnorm/' + k] = np.array(np.sqrt((v.astype(np.float64) ** 2).sum()))
            for k in ('s.0.l.1.weight', 't.11.l.1.weight', 's.5.c.0.weight', 't.7.l.2.weight', 's.3.l.0.weight'):
                fx['gslice/' + k] = grads[k][:8, :8].copy()
        if with_sample:
            Ns = 3
            torch.manual_seed(seed + 3)
            z0s = torch.randn(Ns * B, 45) * 0.8
            torch.manual_seed(seed + 3)
            with torch.no_grad():
                s = model.sample(batch['feat'], N=[Ns, Ns], temp=0.8, mods=['xyz', 'uv', 'verts'], y=y)
            fx['z0_sample'] = npy(z0s)
            for k in ('th_bt', 'logs_t', 'verts', 'xyz', 'uv'):
                fx['sample/' + k] = npy(s[k])
    fx['seed'] = np.array(seed)
    return fx


def mhent_dethead_fixture(mano, B, N, seed):
    """``MHEnt.get_loss`` with the reference's own ``det_head`` (``network.py:376-385``) in the graph: feat -> det_head -> z_det."""
    from mhentropy_b200.synthetic import det_head_state_dict
    model = ref_shim.build_mhent(mano, seed=seed, flow_cfg=PROD)
    model.det_head.load_state_dict(det_head_state_dict(5))
    batch = synthetic_batch(B, N, seed=seed + 1)
    y = {'crop_uv': batch['crop_uv'], 'vis': batch['vis'], 'st': torch.zeros(B, 3), 'image': np.zeros(1)}
    fx = {k: npy(v) for k, v in batch.items() if k != 'z_det'}
    with ref_shim.cpu_mode():
        feat = batch['feat'].clone().requires_grad_(True)
        torch.manual_seed(seed + 2)
        z0 = torch.randn(N * B, 45)
        torch.manual_seed(seed + 2)
        out = model.get_loss(feat, y, mods=['uv'])
        loss = (-out['log_p']).mean()
        loss.backward()
        fx.update({'z0_train': npy(z0), 'log_p': npy(out['log_p']), 'h_q_z_giv_i': npy(out['h_q_z_giv_i']), 'loss': npy(loss),
                   'dfeat': npy(feat.grad), 'z_det': npy(model.det_head(batch['feat']))})
        for k, p in model.det_head.named_parameters():
            g = npy(p.grad)
            fx['gdet/' + k] = g if g.size <= 16 * 512 else g[:16, :64].copy()
            fx['gdetnorm/' + k] = np.array(np.sqrt((g.astype(np.float64) ** 2).sum()))
        for k, p in model.q_z_giv_i.named_parameters():
            fx['gnorm/' + k] = np.array(np.sqrt((npy(p.grad).astype(np.float64) ** 2).sum()))
    fx['seed'] = np.array(seed)
    return fx


def metrics_fixture(mano):
    """Reference ``MHEntLoss.forward`` (``criteria.py:47-173``) run on seeded stand-ins for the outputs of ``MHEnt.sample``."""
    ref_shim.install_stubs(mano)
    with ref_shim.cpu_mode():
        import criteria  # reference hand/criteria.py
    fx = {}
    for tag, N, B, seed in (('a', 7, 6, 21), ('b', 1, 5, 22), ('c', 12, 3, 23)):
        g = torch.Generator().manual_seed(seed)
        pose3d = torch.randn(B, 63, generator=g)
        xyz = pose3d[None] + 0.2 * torch.randn(N, B, 63, generator=g)
        crop_uv = torch.rand(B, 42, generator=g) * 2 - 1
        uv = (crop_uv[None] + 1) / 2 * 256 + 6. * torch.randn(N, B, 42, generator=g)
        scale = 0.05 + 0.1 * torch.rand(B, generator=g)
        vis = (torch.rand(B, 21, generator=g) < 0.7).float()
        if tag == 'a':
            vis[0] = 1.                       # nothing occluded in image 0
            vis[1] = 0.                       # nothing visible in image 1
        if tag == 'c':
            vis[:] = 1.                       # the occluded group is empty for the whole batch (num_valid == 0 branch)
        log_q = torch.randn(N, B, generator=g)
        log_p = torch.randn(B, generator=g)
        target = {'pose3d': pose3d, 'scale': scale, 'crop_uv': crop_uv, 'vis': vis}
        with ref_shim.cpu_mode():
            _, _, metrics = criteria.MHEntLoss()({'log_p': log_p, 'xyz': xyz.clone(), 'uv': uv.clone()}, target)
        for k, v in (('xyz', xyz), ('uv', uv), ('pose3d', pose3d), ('scale', scale), ('crop_uv', crop_uv), ('vis', vis), ('log_q', log_q)):
            fx[f'{tag}/{k}'] = npy(v)
        for k, v in metrics.items():
            fx[f'{tag}/m/{k}'] = npy(v)
        for kk in sorted({1, min(3, N), N}):
            fx[f'{tag}/topk{kk}'] = npy(torch.topk(log_q, kk, dim=0)[1])      # network.py:866-871
    return fx


def main():
    assert ref_shim.reference_available(), 'needs /root/reference'
    mano = synthetic_mano(0)
    if 'dethead' in sys.argv[1:]:
        np.savez_compressed(os.path.join(HERE, 'mhent_dethead.npz'), **mhent_dethead_fixture(mano, B=4, N=10, seed=0))
        print('mhent_dethead.npz', os.path.getsize(os.path.join(HERE, 'mhent_dethead.npz')))
        return
    if 'metrics' in sys.argv[1:]:
        np.savez_compressed(os.path.join(HERE, 'metrics.npz'), **metrics_fixture(mano))
        print('metrics.npz', os.path.getsize(os.path.join(HERE, 'metrics.npz')))
        return
    flows = ref_shim.import_flows(mano)
    np.savez_compressed(os.path.join(HERE, 'flow_small.npz'), **flow_fixture(flows, SMALL, seed=3, R=10, with_weights=True))
    np.savez_compressed(os.path.join(HERE, 'flow_prod.npz'), **flow_fixture(flows, PROD, seed=0, R=12, with_weights=False))
    np.savez_compressed(os.path.join(HERE, 'mano.npz'), **mano_fixture(ref_shim.import_manolayer(mano)))
    np.savez_compressed(os.path.join(HERE, 'mhent_small.npz'),
                        **mhent_fixture(mano, SMALL, B=3, N=10, seed=11, with_weight_grads=True, with_sample=True))
    np.savez_compressed(os.path.join(HERE, 'mhent_prod.npz'),
                        **mhent_fixture(mano, PROD, B=4, N=10, seed=0, with_weight_grads=False, with_sample=False))
    np.savez_compressed(os.path.join(HERE, 'metrics.npz'), **metrics_fixture(mano))
    np.savez_compressed(os.path.join(HERE, 'mhent_dethead.npz'), **mhent_dethead_fixture(mano, B=4, N=10, seed=0))
    for f in sorted(os.listdir(HERE)):
        if f.endswith('.npz'):
            print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == '__main__':
    main()
