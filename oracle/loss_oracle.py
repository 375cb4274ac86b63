"""Oracle (TEST INFRASTRUCTURE): z assembly, reprojection and loss reductions of ``hand/network.py``.

Restates, for the shipped HO3D configuration (``hand/configs/ho3d.yaml``): ``MHEnt._combine_z``
``network.py:703-717``, ``_choose_xyz_from_dec`` ``:466-483`` + ``utils.batch_normalize_pose3d``
``utils.py:46-66``, ``_orth_proj`` ``:497-514`` / ``ManoLayer.batch_orth_proj`` ``ManoLayer.py:150-165``,
``_Laplace.log_prob`` ``:233-258``, ``_ApproxUniform.log_prob`` ``:155-165``, ``_forward_log_p``
``:612-667``, ``_reverse_kld`` ``:760-831``, ``MHEnt.sample`` ``:846-883`` and the loss term of
``MHEntLoss`` ``criteria.py:55,173``.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

from . import flow_oracle as fo
from . import mano_oracle as mo

ROOT_IDX = 12      # network.py:478 ('ho3d')
NORM_IDX = 11      # network.py:479 ('ho3d')
LAPLACE_B = 0.03   # ho3d.yaml:44 b_2d -> network.py:392
TH45_BOX = 2.0     # network.py:427 (use_pca)
TH45_ALPHA = 50.0  # ho3d.yaml:41 w_reg_th
TH3_RADIUS = math.pi  # network.py:431-432
TH3_ALPHA = 5.0
BT_BOX = 0.03      # network.py:433
BT_ALPHA = 50.0
IMAGE_SIZE = 256   # ho3d.yaml:8


def combine_z(x_flow, z_det):
    """``_combine_z``: z = th3 | th45 | bt | logs | t, th45 from the flow, the rest from ``det_head``.

    ``z_det`` columns follow ``zdims`` order restricted to the deterministic entries
    (``network.py:367-373``): th3 (3) | bt (10) | logs (1) | t (2).
    """
    return torch.cat([z_det[:, 0:3], x_flow, z_det[:, 3:13], z_det[:, 13:14], z_det[:, 14:16]], dim=1)


def normalize_pose3d(joints):
    """Root-relative, bone-length-normalised pose — ``utils.py:46-66`` with root 12 / bone 11."""
    root = joints[:, ROOT_IDX:ROOT_IDX + 1]
    rel = joints - root
    bone = torch.sqrt((rel[:, NORM_IDX] ** 2).sum(-1))
    return rel / bone[:, None, None], root, bone


def orth_proj(xyz, logs_t, inv_norm=False):
    """``uv = exp(logs) * xyz[..., :2] + t`` (+ pixel mapping) — ``network.py:497-514``, ``ManoLayer.py:162-165``."""
    s = torch.exp(logs_t[:, 0:1])
    uv = s[:, None, :] * xyz[:, :, :2] + logs_t[:, None, 1:3]
    if inv_norm:
        uv = (uv + 1.0) / 2.0 * IMAGE_SIZE
    return uv


def laplace_log_prob(y, mu, weights, b=LAPLACE_B):
    """``_Laplace.log_prob`` with const b — ``network.py:255-257``."""
    eps = 1e-4
    lp = -(F.relu((y - mu).abs() - eps) + eps) / b - math.log(2 * b)
    return ((weights == 1.0) * lp).flatten(start_dim=1).sum(1)


def box_log_prob(x, a, b, alpha):
    """``_ApproxUniform.log_prob`` sup='rec' — ``network.py:155-158``."""
    return -(alpha * F.relu((x - (a + b) / 2.0).abs() / ((b - a) / 2.0) - 1.0) ** 2).sum(1)


def ball_log_prob(x, radius, alpha):
    """``_ApproxUniform.log_prob`` sup='ball' centred at 0 — ``network.py:159-163``."""
    r = x.norm(p=2, dim=-1)
    return -alpha * F.relu(r / radius - 1.0) ** 2


def th_bt_product(mano_c, z, inv_norm=False):
    """``_th_bt_product`` minus the dead render stub — ``network.py:541-558``."""
    th_bt = z[:, :58]
    dec = mo.mano_wrapper_forward(mano_c, theta=th_bt[:, :48], beta=th_bt[:, -10:])
    xyz, root, bone = normalize_pose3d(dec['mano_joints'])
    verts = (dec['mesh'] - root) / bone[:, None, None]
    uv = orth_proj(xyz, z[:, -3:], inv_norm=inv_norm)
    return {'xyz': xyz, 'verts': verts, 'uv': uv, 'bone': bone, 'dec': dec}


def forward_log_p(mano_c, z, crop_uv, vis, N):
    """``_forward_log_p`` with mods=['uv'] — ``network.py:612-667``.  Returns per-row terms."""
    out = th_bt_product(mano_c, z)
    mu = out['uv'].flatten(start_dim=-2)                                    # (R,42)
    weights = vis[..., None].repeat(N, 1, 2).flatten(start_dim=-2)          # network.py:639-640
    res = {'log_p_uv_giv_z': laplace_log_prob(crop_uv.repeat(N, 1), mu, weights)}
    th3, th45, bt = z[:, :3], z[:, 3:48], z[:, 48:58]
    res['log_p_th3'] = ball_log_prob(th3, TH3_RADIUS, TH3_ALPHA)
    res['log_p_th45'] = box_log_prob(th45, -TH45_BOX, TH45_BOX, TH45_ALPHA)
    res['log_p_bt'] = box_log_prob(bt, -BT_BOX, BT_BOX, BT_ALPHA)
    res['log_p'] = res['log_p_uv_giv_z'] + res['log_p_th3'] + res['log_p_th45'] + res['log_p_bt']   # T = 1
    res['uv'] = out['uv']
    return res


def reverse_kld(sd, mano_c, feat, z_det, z0, crop_uv, vis, N):
    """``MHEnt._reverse_kld`` from ``feat`` on — ``network.py:760-831`` (entropy=True, mods=['uv']).

    Rows are hypothesis-major, r = n*B + b (``feat.repeat(N,1)``, ``network.py:734,747``).
    ``z_det`` (B,16) stands for ``det_head(feat)`` (feature-producer side).
    Two flow passes exactly as the reference: sample (``:733-735``) then ``log_prob`` (``:692``).
    """
    feat_rep = feat.repeat(N, 1)
    x = fo.sample(sd, z0, feat_rep)                                          # network.py:733-735
    z = combine_z(x, z_det.repeat(N, 1))                                     # network.py:747-749
    terms = forward_log_p(mano_c, z, crop_uv, vis, N)
    q_log_p = terms['log_p'].reshape(N, -1).mean(0)                          # network.py:793
    log_q = fo.log_prob(sd, z[:, 3:48], feat_rep)                            # network.py:801 -> flows.py:271
    h = (-log_q).reshape(N, -1).mean(0)                                      # network.py:802
    log_p = h + q_log_p                                                      # network.py:803-808
    return {
        'log_p': log_p, 'h_q_z_giv_i': h, 'q_log_p_z_giv_y': q_log_p, 'z': z, 'x': x, 'log_q': log_q,
        'uv': terms['uv'], 'row_log_p': terms['log_p'],
        'th_norm': z[:, :48].norm(p=2, dim=1), 'bt_norm': z[:, 48:58].norm(p=2, dim=1),
    }


def mhent_loss(log_p):
    """``MHEntLoss``: mean_B(-log_p) — ``criteria.py:55,173``."""
    return (-log_p).mean()


def mhent_sample(sd, mano_c, feat, z_det, z0, N):
    """``MHEnt.sample`` with N_quant == N, mods={xyz,uv,verts} — ``network.py:846-883``."""
    B = feat.shape[0]
    x = fo.sample(sd, z0, feat.repeat(N, 1))
    z = combine_z(x, z_det.repeat(N, 1))
    out = th_bt_product(mano_c, z, inv_norm=True)
    return {
        'th_bt': z[:, :58].reshape(N, B, 58), 'logs_t': z[:, -3:].reshape(N, B, 3),
        'verts': out['verts'].reshape(N, B, -1), 'xyz': out['xyz'].reshape(N, B, -1),
        'uv': out['uv'].reshape(N, B, -1),
    }


from mhentropy_b200.synthetic import synthetic_batch  # noqa: E402,F401  (shared recipe)
