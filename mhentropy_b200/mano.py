"""Drop-in MANO layers on the B200 kernels.

``ManoCore`` mirrors ``manopth.manolayer.ManoLayer`` (reference ``hand/manopth/manolayer.py:13-274``):
same buffer names (``th_shapedirs``, ``th_posedirs``, ``th_v_template``, ``th_J_regressor``,
``th_weights``, ``th_faces``, ``th_hands_mean``, ``th_comps``, ``th_selected_comps``, ``th_betas``),
``forward(th_pose_coeffs, th_betas) -> (verts, jtr)`` in millimetres.

``ManoLayer`` mirrors the wrapper ``hand/ManoLayer.py:10-165``: ``forward(z=None, beta=None,
theta=None) -> dict(beta, theta, mesh, joints, mano_joints)``, ``.mano_layer``, ``.Jreg``,
``.mano_faces``, static ``batch_orth_proj``.  The dead ``render`` stub (``ManoLayer.py:62-105``, its
renderer is commented out upstream) is not reproduced.

Accelerated configuration: ``use_pca=True, ncomps=45``, axis-angle root, right hand, ``center_idx=9``
(``ManoLayer.py:19-21``, ``CrossModalHand.py:72-74``).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn

from . import _lib
from ._lib import ManoConsts, check, lib, ptr, stream_ptr
from .mano_assets import resolve_mano

FREIHAND2RHD = [0, 4, 3, 2, 1, 8, 7, 6, 5, 12, 11, 10, 9, 16, 15, 14, 13, 20, 19, 18, 17]          # utils.py:15
RHD2BIGHAND = [0, 4, 8, 12, 16, 20, 3, 2, 1, 7, 6, 5, 11, 10, 9, 15, 14, 13, 19, 18, 17]           # utils.py:12


class _ManoFn(torch.autograd.Function):
    """(theta, beta) -> (verts, jtr, joints2); any output may be skipped."""

    @staticmethod
    def forward(ctx, theta, beta, core, order, want_verts, want_joints2):
        theta = theta.contiguous().float()
        beta = beta.contiguous().float()
        _lib.require_cuda_f32(theta, beta)
        R = theta.shape[0]
        dev = theta.device
        consts = core._consts(dev)
        verts = torch.empty(R, 778, 3, device=dev) if (want_verts or want_joints2) else None
        jtr = torch.empty(R, 21, 3, device=dev)
        joints2 = torch.empty(R, 21, 3, device=dev) if want_joints2 else None
        wsb = lib().mhe_mano_workspace_bytes(R, 0)
        ws = _lib.WORKSPACE.get(wsb, dev)
        check(lib().mhe_mano_fwd(consts, ptr(theta), theta.shape[1], ptr(beta), beta.shape[1], R, order, ptr(verts), ptr(jtr),
                                 ptr(joints2), ptr(ws), wsb, stream_ptr(dev)), 'mhe_mano_fwd')
        ctx.save_for_backward(theta, beta)
        ctx.core, ctx.order = core, order
        outs = (verts if verts is not None else theta.new_empty(0), jtr, joints2 if joints2 is not None else theta.new_empty(0))
        if verts is None:
            ctx.mark_non_differentiable(outs[0])
        if joints2 is None:
            ctx.mark_non_differentiable(outs[2])
        ctx.has = (verts is not None, joints2 is not None)
        return outs

    @staticmethod
    def backward(ctx, dverts, djtr, dj2):
        theta, beta = ctx.saved_tensors
        R = theta.shape[0]
        dev = theta.device
        consts = ctx.core._consts(dev)
        dverts = dverts.contiguous() if (ctx.has[0] and dverts is not None) else None
        dj2 = dj2.contiguous() if (ctx.has[1] and dj2 is not None) else None
        djtr = djtr.contiguous() if djtr is not None else None
        dtheta = torch.empty_like(theta)
        dbeta = torch.empty_like(beta)
        mesh = int(dverts is not None or dj2 is not None)
        wsb = lib().mhe_mano_workspace_bytes(R, mesh)
        ws = _lib.WORKSPACE.get(wsb, dev)
        check(lib().mhe_mano_bwd(consts, ptr(theta), theta.shape[1], ptr(beta), beta.shape[1], R, ctx.order, ptr(dverts),
                                 ptr(djtr), ptr(dj2), ptr(dtheta), theta.shape[1], ptr(dbeta), beta.shape[1], 0, ptr(ws), wsb,
                                 stream_ptr(dev)), 'mhe_mano_bwd')
        return dtheta, dbeta, None, None, None, None


class ManoCore(nn.Module):
    """``manopth`` ManoLayer equivalent (reference ``manolayer.py:13-274``)."""

    def __init__(self, center_idx=9, flat_hand_mean=False, ncomps=45, side='right', mano_root='mano/models', use_pca=True,
                 root_rot_mode='axisang', joint_rot_mode='axisang', robust_rot=False, mano_data: dict | None = None,
                 synthetic_seed: int | None = None):
        super().__init__()
        if not (use_pca and ncomps == 45 and side == 'right' and root_rot_mode == 'axisang' and center_idx == 9):
            raise NotImplementedError('accelerated MANO path: use_pca=True, ncomps=45, right hand, axis-angle root, center_idx=9')
        self.center_idx, self.ncomps, self.side, self.use_pca = center_idx, ncomps, side, use_pca
        self.rot = 3
        self.flat_hand_mean = flat_hand_mean
        if mano_data is None:
            mano_data, self.mano_source = resolve_mano(mano_root, synthetic_seed)
        else:
            self.mano_source = 'provided'
        f32 = lambda a: torch.as_tensor(np.asarray(a, dtype=np.float64)).float()  # noqa: E731
        comps = np.asarray(mano_data['hands_components'])
        hands_mean = np.zeros(comps.shape[1]) if flat_hand_mean else np.asarray(mano_data['hands_mean'])
        self.register_buffer('th_betas', f32(mano_data['betas']).unsqueeze(0))
        self.register_buffer('th_shapedirs', f32(mano_data['shapedirs']))
        self.register_buffer('th_posedirs', f32(mano_data['posedirs']))
        self.register_buffer('th_v_template', f32(mano_data['v_template']).unsqueeze(0))
        self.register_buffer('th_J_regressor', f32(mano_data['J_regressor']))
        self.register_buffer('th_weights', f32(mano_data['weights']))
        self.register_buffer('th_faces', torch.as_tensor(np.asarray(mano_data['f']).astype(np.int32)).long())
        self.register_buffer('th_hands_mean', f32(hands_mean).unsqueeze(0))
        self.register_buffer('th_comps', f32(comps))
        self.register_buffer('th_selected_comps', f32(comps[:ncomps]))
        self.kintree_parents = list(np.asarray(mano_data['kintree_table'])[0].tolist())
        self._packed = None

    def _consts(self, device) -> ManoConsts:
        """Pack the buffers into the kernel layouts (once per device; fp64 products for jt / js)."""
        # derived copies are keyed on the device AND on the buffers' identity / version counters: load_state_dict copies into the
        # th_* buffers in place, and a stale jt / js / posedirs plane set would silently mix two hands
        srcs = (self.th_selected_comps, self.th_hands_mean, self.th_v_template, self.th_shapedirs, self.th_posedirs,
                self.th_J_regressor, self.th_weights)
        key = (str(device),) + tuple((t.data_ptr(), t._version) for t in srcs)
        if self._packed is None or self._packed[0] != key:
            d = lambda t: t.detach().double().cpu()  # noqa: E731
            jreg = d(self.th_J_regressor)
            pack = {
                'comps': self.th_selected_comps, 'hands_mean': self.th_hands_mean.reshape(-1),
                'v_template': self.th_v_template.reshape(778, 3), 'shapedirs': self.th_shapedirs,
                'posedirs_t': self.th_posedirs.reshape(778 * 3, 135).t(), 'jreg': self.th_J_regressor,
                'weights': self.th_weights,
                'jt': (jreg @ d(self.th_v_template).reshape(778, 3)).float(),
                'js': torch.einsum('jv,vck->jck', jreg, d(self.th_shapedirs)).float(),
            }
            tensors = {k: v.detach().to(device=device, dtype=torch.float32).contiguous() for k, v in pack.items()}
            c = ManoConsts(**{k: v.data_ptr() for k, v in tensors.items()}, pose_tables=None, posedirs_planes=None)
            if torch.device(device).type == 'cuda':   # gather the small pose / tip tables once (coalesced staging in the kernels)
                tensors['pose_tables'] = torch.empty(lib().mhe_mano_pose_tables_floats(), device=device, dtype=torch.float32)
                check(lib().mhe_mano_pack_pose_tables(c, ptr(tensors['pose_tables']), stream_ptr(torch.device(device))), 'mhe_mano_pack_pose_tables')
                c.pose_tables = tensors['pose_tables'].data_ptr()
                # posedirs as split half planes: the pose blend of the mesh forward runs as one tcgen05 GEMM over all rows
                tensors['posedirs_planes'] = torch.empty(lib().mhe_mano_posedirs_planes_bytes(), device=device, dtype=torch.uint8)
                check(lib().mhe_mano_pack_posedirs_planes(c, ptr(tensors['posedirs_planes']), stream_ptr(torch.device(device))),
                      'mhe_mano_pack_posedirs_planes')
                c.posedirs_planes = tensors['posedirs_planes'].data_ptr()
            self._packed = (key, tensors, c)
        return self._packed[2]

    def _apply(self, fn, *args, **kwargs):
        self._packed = None
        return super()._apply(fn, *args, **kwargs)

    def _load_from_state_dict(self, *args, **kwargs):
        super()._load_from_state_dict(*args, **kwargs)
        self._packed = None                  # the kernels' derived constants follow the loaded buffers

    def forward(self, th_pose_coeffs, th_betas=None, th_trans=None, root_palm=None, share_betas=None):
        if th_trans is not None and bool(torch.as_tensor(th_trans).abs().sum() != 0):
            raise NotImplementedError('th_trans is outside the accelerated path')
        if th_betas is None or th_betas.numel() == 1:
            th_betas = self.th_betas.to(th_pose_coeffs.device).expand(th_pose_coeffs.shape[0], 10)
        verts, jtr, _ = _ManoFn.apply(th_pose_coeffs, th_betas, self, 0, True, False)
        return verts, jtr


class ManoLayer(nn.Module):
    """Wrapper equivalent to reference ``hand/ManoLayer.py:10-165``."""

    def __init__(self, MANO_dir='./mano/', flat_hand_mean=True, ncomps=45, use_pca=False, n_latent=None, skeidx='FreiHand',
                 output_size=256, mask_sz=256, mano_data: dict | None = None, synthetic_seed: int | None = None):
        super().__init__()
        self.mano_layer = ManoCore(center_idx=9, flat_hand_mean=flat_hand_mean, ncomps=ncomps, side='right', mano_root=MANO_dir,
                                   use_pca=use_pca, mano_data=mano_data, synthetic_seed=synthetic_seed)
        self.Jreg = self.mano_layer.th_J_regressor
        self.n_latent = n_latent
        if n_latent is not None:
            self.mano_beta = nn.Sequential(nn.Linear(n_latent, 512), nn.ReLU(), nn.Linear(512, 10))
            self.mano_theta = nn.Sequential(nn.Linear(n_latent, 512), nn.ReLU(), nn.Linear(512, 48))
        self.skeidx = skeidx
        self.output_size = output_size
        self.mask_sz = mask_sz
        self.mano_faces = self.mano_layer.th_faces
        self.f = self.mano_faces[None, :, :].int()

    def forward(self, z=None, beta=None, theta=None, want_mesh=True, want_joints=True):
        """``want_mesh`` / ``want_joints`` (extensions, default = reference behaviour) let a caller that only
        consumes ``mano_joints`` skip the 778-vertex mesh and the regressed second joint set."""
        if beta is None:
            beta = self.mano_beta(z)
        if theta is None:
            theta = self.mano_theta(z)
        beta = beta.reshape(-1, 10)
        theta = theta.reshape(-1, 48)
        order = 0 if self.skeidx == 'FreiHand' else 1
        verts, mano_joints, joints = _ManoFn.apply(theta, beta, self.mano_layer, order, want_mesh, want_mesh and want_joints)
        if self.skeidx == 'BigHand':
            joints = joints[:, RHD2BIGHAND, :]
            mano_joints = mano_joints[:, RHD2BIGHAND, :]
        return {'beta': beta, 'theta': theta, 'mesh': verts if want_mesh else None,
                'joints': joints if (want_mesh and want_joints) else None, 'mano_joints': mano_joints}

    @staticmethod
    def batch_orth_proj(joint, scale_camera, trans_camera, image_size: int = 256, inv_norm=True):
        """``ManoLayer.py:150-165``: uv = s * xyz[..., :2] + t (then to pixels)."""
        out = scale_camera[:, None, :] * joint[:, :, :2] + trans_camera[:, None, :]
        if inv_norm:
            out = (out + 1.) / 2. * image_size
        return out
