"""Oracle (TEST INFRASTRUCTURE): MANO layer of reference ``hand/manopth/manolayer.py`` + wrapper.

Configured path only: ``use_pca=True, ncomps=45, flat_hand_mean=False,
center_idx=9, side='right', root_rot_mode='axisang'`` (reference ``ManoLayer.py:19-21``,
``CrossModalHand.py:72-74``).  Written from the maths — a parent-indexed kinematic chain
instead of the reference's three hard-coded levels — and checked against the reference in
``tests/test_oracle_vs_golden.py``.
"""
from __future__ import annotations

import torch

from mhentropy_b200.mano_assets import KINTREE_PARENTS

# reference manolayer.py:250 (right hand) — tip vertices appended as joints 16..20
TIP_VERTS = [745, 317, 444, 556, 673]
# reference manolayer.py:260 — MANO(16)+tips(5) -> visualisation order
JOINT_REORDER = [0, 13, 14, 15, 16, 1, 2, 3, 17, 4, 5, 6, 18, 10, 11, 12, 19, 7, 8, 9, 20]
# reference utils.py:15 — FreiHand -> RHD skeleton order (wrapper ManoLayer.py:54-56)
FREIHAND2RHD = [0, 4, 3, 2, 1, 8, 7, 6, 5, 12, 11, 10, 9, 16, 15, 14, 13, 20, 19, 18, 17]
# wrapper ManoLayer.py:113-127 — second joint set: tip vertices and regressed-joint placement
WRAP_TIP_VERTS = {4: 744, 8: 320, 12: 443, 16: 555, 20: 672}
WRAP_MAPPING = {0: 0, 1: 5, 2: 6, 3: 7, 4: 9, 5: 10, 6: 11, 7: 17, 8: 18, 9: 19,
                10: 13, 11: 14, 12: 15, 13: 1, 14: 2, 15: 3}
CENTER_IDX = 9


def mano_constants(mano: dict, dtype=torch.float32) -> dict:
    """The buffers ``manolayer.py:71-99`` registers, as tensors of ``dtype``."""
    t = lambda a: torch.as_tensor(a, dtype=torch.float64).to(dtype)  # noqa: E731
    return {
        'shapedirs': t(mano['shapedirs']),           # (778,3,10)
        'posedirs': t(mano['posedirs']),             # (778,3,135)
        'v_template': t(mano['v_template']),         # (778,3)
        'J_regressor': t(mano['J_regressor']),       # (16,778)
        'weights': t(mano['weights']),               # (778,16)
        'hands_mean': t(mano['hands_mean']),         # (45,)
        'comps': t(mano['hands_components']),        # (45,45)  (ncomps=45 -> all rows selected)
        'faces': torch.as_tensor(mano['f'].astype('int64')),
    }


def rodrigues(axisang):
    """Axis-angle (N,3) -> rotation matrices (N,3,3) via the half-angle quaternion.

    Reference ``rodrigues_layer.py:43-54`` (norm taken of ``v + 1e-8``) and ``quat2mat`` ``:15-40``
    (quaternion re-normalised before use).
    """
    angle = torch.norm(axisang + 1e-8, p=2, dim=1, keepdim=True)
    axis = axisang / angle
    half = angle * 0.5
    quat = torch.cat([torch.cos(half), torch.sin(half) * axis], dim=1)
    quat = quat / quat.norm(p=2, dim=1, keepdim=True)
    w, x, y, z = quat[:, 0], quat[:, 1], quat[:, 2], quat[:, 3]
    w2, x2, y2, z2 = w * w, x * x, y * y, z * z
    wx, wy, wz, xy, xz, yz = w * x, w * y, w * z, x * y, x * z, y * z
    rot = torch.stack([
        w2 + x2 - y2 - z2, 2 * xy - 2 * wz, 2 * wy + 2 * xz,
        2 * wz + 2 * xy, w2 - x2 + y2 - z2, 2 * yz - 2 * wx,
        2 * xz - 2 * wy, 2 * wx + 2 * yz, w2 - x2 - y2 + z2], dim=1)
    return rot.view(-1, 3, 3)


def mano_forward(c: dict, theta, beta):
    """``manopth.ManoLayer.forward(theta (R,48), beta (R,10)) -> verts (R,778,3) mm, jtr (R,21,3) mm``.

    Reference ``manolayer.py:110-274`` (steps numbered as in SURVEY.md §3.4).
    """
    R = theta.shape[0]
    # 1. PCA coefficients -> axis-angle, add the mean pose (manolayer.py:131-143)
    hand = theta[:, 3:48] @ c['comps'] + c['hands_mean']
    full_pose = torch.cat([theta[:, :3], hand], dim=1)                     # (R,48)
    # 2. 16 rotations; pose feature = R - I of the 15 articulated joints (tensutils.py:6-12)
    rots = rodrigues(full_pose.reshape(-1, 3)).view(R, 16, 3, 3)
    eye = torch.eye(3, dtype=theta.dtype, device=theta.device)
    pose_map = (rots[:, 1:] - eye).reshape(R, 135)
    # 3. shape blend + joint regression (manolayer.py:181-184)
    v_shaped = torch.einsum('vdk,rk->rvd', c['shapedirs'], beta) + c['v_template']
    J = torch.einsum('jv,rvd->rjd', c['J_regressor'], v_shaped)            # (R,16,3)
    # 4. pose blend (manolayer.py:187-188)
    v_posed = v_shaped + torch.einsum('vdk,rk->rvd', c['posedirs'], pose_map)
    # 5. kinematic chain (manolayer.py:193-229); G_k = G_parent * [R_k | J_k - J_parent]
    G_rot = [None] * 16
    G_tr = [None] * 16
    G_rot[0] = rots[:, 0]
    G_tr[0] = J[:, 0]
    for k in range(1, 16):
        p = KINTREE_PARENTS[k]
        G_rot[k] = G_rot[p] @ rots[:, k]
        G_tr[k] = (G_rot[p] @ (J[:, k] - J[:, p]).unsqueeze(-1)).squeeze(-1) + G_tr[p]
    G_rot = torch.stack(G_rot, dim=1)                                       # (R,16,3,3)
    G_tr = torch.stack(G_tr, dim=1)                                         # (R,16,3)
    # 6. remove the rest-pose joint location, blend per vertex, apply (manolayer.py:232-246)
    A_tr = G_tr - (G_rot @ J.unsqueeze(-1)).squeeze(-1)
    T_rot = torch.einsum('vk,rkij->rvij', c['weights'], G_rot)
    T_tr = torch.einsum('vk,rki->rvi', c['weights'], A_tr)
    verts = (T_rot @ v_posed.unsqueeze(-1)).squeeze(-1) + T_tr
    # 7. joints = chain translations + 5 tip vertices, reorder, centre on joint 9, metres -> mm
    jtr = torch.cat([G_tr, verts[:, TIP_VERTS]], dim=1)[:, JOINT_REORDER]
    center = jtr[:, CENTER_IDX:CENTER_IDX + 1]
    jtr = jtr - center
    verts = verts - center
    return verts * 1000, jtr * 1000


def xyz_from_vertice(c: dict, verts):
    """Wrapper's second joint set, FreiHand order — reference ``ManoLayer.py:109-148``."""
    reg = torch.einsum('jv,rvd->rjd', c['J_regressor'].to(verts.dtype), verts)   # (R,16,3)
    out = [None] * 21
    for mano_id, my_id in WRAP_MAPPING.items():
        out[my_id] = reg[:, mano_id]
    for my_id, vid in WRAP_TIP_VERTS.items():
        out[my_id] = verts[:, vid]
    return torch.stack(out, dim=1)                                           # (R,21,3)


def mano_wrapper_forward(c: dict, theta, beta, skeidx='RHD'):
    """``hand/ManoLayer.py:45-60`` with ``skeidx='RHD'`` (``network.py:360-363``)."""
    beta = beta.reshape(-1, 10)
    theta = theta.reshape(-1, 48)
    verts, mano_joints = mano_forward(c, theta, beta)
    joints = xyz_from_vertice(c, verts)
    if skeidx == 'RHD':
        joints = joints[:, FREIHAND2RHD]
        mano_joints = mano_joints[:, FREIHAND2RHD]
    return {'beta': beta, 'theta': theta, 'mesh': verts, 'joints': joints, 'mano_joints': mano_joints}
