"""The per-row math headers of the CUDA kernels (csrc/mano_math.cuh, csrc/loss_math.cuh), compiled for the
host with g++, against the oracle and its autograd.  Covers the hand-derived gradients without a GPU."""
import ctypes
import os
import subprocess

import numpy as np
import pytest
import torch

from mhentropy_b200._lib import LossCfg
from mhentropy_b200.losses import default_loss_cfg
from mhentropy_b200.mano_assets import KINTREE_PARENTS, synthetic_mano
from oracle import loss_oracle as lo
from oracle import mano_oracle as mo

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
F = ctypes.POINTER(ctypes.c_float)


def fp(a):
    assert a.dtype == np.float32 and a.flags['C_CONTIGUOUS']
    return a.ctypes.data_as(F)


@pytest.fixture(scope='module')
def hm(tmp_path_factory):
    out = str(tmp_path_factory.mktemp('hm') / 'libhm.so')
    subprocess.check_call(['g++', '-O1', '-std=c++17', '-shared', '-fPIC',
                           os.path.join(ROOT, 'tests', 'hostmath', 'hostmath.cpp'), '-o', out])
    lib = ctypes.CDLL(out)
    lib.hm_row_log_p.restype = ctypes.c_float
    lib.hm_row_log_p_bwd.argtypes = [ctypes.POINTER(LossCfg), F, F, F, F, ctypes.c_float, F, F]
    return lib


def test_rodrigues_forward_backward(hm):
    rs = np.random.RandomState(0)
    vs = [rs.normal(size=3) * s for s in (1.5, 0.3, 1e-3, 3.0)] + [np.zeros(3)]
    for v in vs:
        v32 = v.astype(np.float32)
        R = np.zeros(9, np.float32)
        hm.hm_rodrigues_fwd(fp(v32), fp(R))
        vt = torch.tensor(v32, dtype=torch.float64).view(1, 3).requires_grad_(True)
        Rt = mo.rodrigues(vt).reshape(9)
        assert np.abs(R - Rt.detach().numpy()).max() < 2e-6
        g = rs.normal(size=9).astype(np.float32)
        (Rt * torch.tensor(g, dtype=torch.float64)).sum().backward()
        dv = np.zeros(3, np.float32)
        hm.hm_rodrigues_bwd(fp(v32), fp(g), fp(dv))
        assert np.abs(dv - vt.grad.numpy().reshape(3)).max() < 5e-5 * max(1.0, np.abs(vt.grad.numpy()).max())


def _oracle_pose(c, theta, beta):
    """A [16][12], Gt [16][3], pm [135] from the oracle's formulas (fp64)."""
    hand = theta[3:48] @ c['comps'] + c['hands_mean']
    full = torch.cat([theta[:3], hand])
    rots = mo.rodrigues(full.reshape(-1, 3))
    v_shaped = torch.einsum('vdk,k->vd', c['shapedirs'], beta) + c['v_template']
    J = c['J_regressor'] @ v_shaped
    Gr, Gt = [rots[0]], [J[0]]
    for k in range(1, 16):
        p = KINTREE_PARENTS[k]
        Gr.append(Gr[p] @ rots[k])
        Gt.append(Gr[p] @ (J[k] - J[p]) + Gt[p])
    Gr, Gt = torch.stack(Gr), torch.stack(Gt)
    A = torch.cat([Gr.reshape(16, 9), Gt - (Gr @ J.unsqueeze(-1)).squeeze(-1)], dim=1)
    pm = (rots[1:] - torch.eye(3, dtype=theta.dtype)).reshape(135)
    return A, Gt, pm


def test_pose_chain_forward_backward(hm):
    mano = synthetic_mano(0)
    c = mo.mano_constants(mano, torch.float64)
    jreg = c['J_regressor']
    jt = (jreg @ c['v_template']).float().numpy().copy()
    js = torch.einsum('jv,vck->jck', jreg, c['shapedirs']).float().numpy().copy()
    comps = c['comps'].float().numpy().copy()
    mean = c['hands_mean'].float().numpy().copy()
    rs = np.random.RandomState(1)
    for trial in range(3):
        theta = rs.normal(size=48).astype(np.float32) * 0.8
        beta = rs.normal(size=10).astype(np.float32) * 0.03
        A = np.zeros(192, np.float32); Gt = np.zeros(48, np.float32); pm = np.zeros(135, np.float32)
        hm.hm_pose_fwd(fp(comps), fp(mean), fp(jt), fp(js), fp(theta), fp(beta), fp(A), fp(Gt), fp(pm))
        th = torch.tensor(theta, dtype=torch.float64, requires_grad=True)
        bt = torch.tensor(beta, dtype=torch.float64, requires_grad=True)
        Ao, Gto, pmo = _oracle_pose(c, th, bt)
        assert np.abs(A - Ao.detach().numpy().reshape(-1)).max() < 5e-6
        assert np.abs(Gt - Gto.detach().numpy().reshape(-1)).max() < 5e-6
        assert np.abs(pm - pmo.detach().numpy()).max() < 5e-6
        dA = rs.normal(size=192).astype(np.float32); dGt = rs.normal(size=48).astype(np.float32); dpm = rs.normal(size=135).astype(np.float32)
        ((Ao.reshape(-1) * torch.tensor(dA, dtype=torch.float64)).sum() + (Gto.reshape(-1) * torch.tensor(dGt, dtype=torch.float64)).sum()
         + (pmo * torch.tensor(dpm, dtype=torch.float64)).sum()).backward()
        dth = np.zeros(48, np.float32); dbt = np.zeros(10, np.float32)
        hm.hm_pose_bwd(fp(comps), fp(mean), fp(jt), fp(js), fp(theta), fp(beta), fp(dGt), fp(dA), fp(dpm), fp(dth), fp(dbt))
        assert np.abs(dth - th.grad.numpy()).max() < 2e-4 * np.abs(th.grad.numpy()).max()
        assert np.abs(dbt - bt.grad.numpy()).max() < 2e-4 * np.abs(bt.grad.numpy()).max()


def test_row_log_p_forward_backward(hm):
    cfg = default_loss_cfg()
    rs = np.random.RandomState(2)
    for trial in range(4):
        j = (rs.normal(size=(21, 3)) * 60).astype(np.float32)
        z = np.concatenate([rs.normal(size=3) * (2.5 if trial % 2 else 0.5), rs.normal(size=45) * 1.5, rs.normal(size=10) * 0.03,
                            [np.log(0.3) + 0.1 * rs.normal()], rs.normal(size=2) * 0.1]).astype(np.float32)
        crop = rs.uniform(-1, 1, size=42).astype(np.float32)
        vis = (rs.uniform(size=21) < 0.7).astype(np.float32)
        uv = np.zeros(42, np.float32)
        lp = hm.hm_row_log_p(ctypes.byref(cfg), fp(j.reshape(-1).copy()), fp(z), fp(crop), fp(vis), fp(uv))
        jt_ = torch.tensor(j, dtype=torch.float64).unsqueeze(0).requires_grad_(True)
        zt = torch.tensor(z, dtype=torch.float64).unsqueeze(0).requires_grad_(True)
        xyz, _, _ = lo.normalize_pose3d(jt_)
        uvo = lo.orth_proj(xyz, zt[:, -3:])
        w = torch.tensor(vis, dtype=torch.float64)[None, :, None].repeat(1, 1, 2).flatten(1)
        lpo = (lo.laplace_log_prob(torch.tensor(crop, dtype=torch.float64)[None], uvo.flatten(1), w)
               + lo.ball_log_prob(zt[:, :3], lo.TH3_RADIUS, lo.TH3_ALPHA) + lo.box_log_prob(zt[:, 3:48], -2.0, 2.0, 50.0)
               + lo.box_log_prob(zt[:, 48:58], -0.03, 0.03, 50.0))
        assert abs(lp - lpo.item()) < 2e-5 * abs(lpo.item())
        assert np.abs(uv - uvo.detach().numpy().reshape(-1)).max() < 1e-5
        lpo.backward()
        dj = np.zeros(63, np.float32); dz = np.zeros(61, np.float32)
        hm.hm_row_log_p_bwd(ctypes.byref(cfg), fp(j.reshape(-1).copy()), fp(z), fp(crop), fp(vis), ctypes.c_float(1.0), fp(dj), fp(dz))
        assert np.abs(dj - jt_.grad.numpy().reshape(-1)).max() < 2e-4 * np.abs(jt_.grad.numpy()).max()
        assert np.abs(dz - zt.grad.numpy().reshape(-1)).max() < 2e-4 * np.abs(zt.grad.numpy()).max()
