"""Reprojection / likelihood / prior / entropy reductions of ``MHEnt`` on the B200 kernels, and
``MHEntHead`` — the part of reference ``hand/network.py:MHEnt`` downstream of the image feature
(``_sample_q_z_giv_i`` :719, ``_forward_log_p`` :612, ``_reverse_log_q`` :669, ``_reverse_kld`` :760,
``get_loss`` :838, ``sample`` :846), with the same method names and return keys.
"""
from __future__ import annotations

import math
import weakref

import torch
import torch.nn as nn

from . import _lib
from ._lib import LossCfg, check, lib, ptr, stream_ptr
from .flows import RealNVP
from .mano import ManoLayer


def default_loss_cfg() -> LossCfg:
    """HO3D configuration: b_2d 0.03 (ho3d.yaml:44), th45 box 2 / alpha 50 (network.py:427, ho3d.yaml:41),
    th3 ball pi / alpha 5 (network.py:431-432), beta box 0.03 / alpha 50 (network.py:433), root 12 / bone 11
    (network.py:478-479)."""
    return LossCfg(0.03, 2.0, 50.0, math.pi, 5.0, 0.03, 50.0, 12, 11)


class _CombineZFn(torch.autograd.Function):
    """z = th3 | th45(flow) | bt | logs | t  (``network.py:703-717``), z_det repeated over hypotheses."""

    @staticmethod
    def forward(ctx, x_flow, z_det):
        x_flow = x_flow.contiguous().float()
        z_det = z_det.contiguous().float()
        _lib.require_cuda_f32(x_flow, z_det)
        R, B = x_flow.shape[0], z_det.shape[0]
        z = torch.empty(R, 61, device=x_flow.device)
        check(lib().mhe_combine_z_fwd(ptr(x_flow), ptr(z_det), R, B, ptr(z), stream_ptr(z.device)), 'mhe_combine_z_fwd')
        ctx.dims = (R, B)
        return z

    @staticmethod
    def backward(ctx, dz):
        R, B = ctx.dims
        dz = dz.contiguous()
        dx = torch.empty(R, 45, device=dz.device)
        dzd = torch.empty(B, 16, device=dz.device)
        check(lib().mhe_combine_z_bwd(ptr(dz), R, B, ptr(dx), ptr(dzd), stream_ptr(dz.device)), 'mhe_combine_z_bwd')
        return dx, dzd


class _ReprojLossFn(torch.autograd.Function):
    """(joints, z, log_q | crop_uv, vis) -> log_p (B,), h (B,), q_log_p (B,), uv (R,42), row_log_p (R,).
    Differentiable through ``log_p`` only."""

    @staticmethod
    def forward(ctx, joints, z, log_q, crop_uv, vis, cfg):
        joints = joints.contiguous().float()
        z = z.contiguous().float()
        log_q = log_q.contiguous().float()
        crop_uv = crop_uv.contiguous().float()
        vis = vis.contiguous().float()
        _lib.require_cuda_f32(joints, z, log_q, crop_uv, vis)
        R, B = z.shape[0], crop_uv.shape[0]
        dev = z.device
        uv = torch.empty(R, 42, device=dev)
        row = torch.empty(R, device=dev)
        log_p = torch.empty(B, device=dev)
        h = torch.empty(B, device=dev)
        qlp = torch.empty(B, device=dev)
        check(lib().mhe_reproj_loss_fwd(cfg, ptr(joints), ptr(z), ptr(crop_uv), ptr(vis), ptr(log_q), R, B, ptr(uv), ptr(row),
                                        ptr(log_p), ptr(h), ptr(qlp), None, stream_ptr(dev)), 'mhe_reproj_loss_fwd')
        ctx.save_for_backward(joints, z, crop_uv, vis)
        ctx.cfg = cfg
        ctx.mark_non_differentiable(h, qlp, uv, row)
        return log_p, h, qlp, uv, row

    @staticmethod
    def backward(ctx, dlog_p, *_):
        joints, z, crop_uv, vis = ctx.saved_tensors
        R, B = z.shape[0], crop_uv.shape[0]
        dev = z.device
        dlog_p = dlog_p.contiguous()
        dj = torch.empty_like(joints)
        dz = torch.empty_like(z)
        dlq = torch.empty(R, device=dev)
        check(lib().mhe_reproj_loss_bwd(ctx.cfg, ptr(joints), ptr(z), ptr(crop_uv), ptr(vis), R, B, ptr(dlog_p), None, ptr(dj),
                                        ptr(dz), ptr(dlq), stream_ptr(dev)), 'mhe_reproj_loss_bwd')
        return dj, dz, dlq, None, None, None


def _release_engine(eng, token):
    if getattr(eng, 'token', None) is token:          # the graph was dropped without a backward: the engine is free again
        eng.busy, eng.token = False, None


class _FusedLossFn(torch.autograd.Function):
    """``MHEnt._reverse_kld`` from the feature on (``network.py:760-831``) as ONE autograd node: forward and backward are two captured
    CUDA graphs of the engine (``engine.TrainStep`` split mode), so a ``get_loss`` + ``backward()`` costs two graph launches on the host
    instead of ~25 kernel launches through eight autograd functions.  Differentiable through ``log_p`` (B,) with ANY downstream loss:
    the backward graph takes dL/dlog_p per image.  Parameter gradients land in the flow's flat gradient buffer (``p.grad`` are views)."""

    @staticmethod
    def forward(ctx, feat, z_det, anchor, head, eng, z0, crop_uv, vis):
        flow = head.q_z_giv_i
        flow.packed_weights(eng.dev)                  # re-packs the split planes here (eagerly, stream-ordered) if the weights changed
        eng.load(feat.detach(), z_det.detach(), z0, crop_uv, vis)
        eng.forward_graph().replay()
        outs = tuple(t.clone() for t in (eng.log_p, eng.h, eng.qlp, eng.uv, eng.norms[0], eng.norms[1]))
        ctx.head, ctx.eng = head, eng
        if any(ctx.needs_input_grad[:3]):             # the engine's buffers hold this forward until its backward (or until the graph is dropped)
            eng.busy = True
            eng.token = token = object()
            ctx.token = token
            ctx._fin = weakref.finalize(ctx, _release_engine, eng, token)
        ctx.mark_non_differentiable(*outs[1:])
        return outs

    @staticmethod
    def backward(ctx, dlog_p, *_):
        eng, flow = ctx.eng, ctx.head.q_z_giv_i
        if getattr(eng, 'token', None) is not ctx.token:
            raise RuntimeError('the fused loss node was run again before this backward (retain_graph / double backward are not supported on '
                               'the fused path: set head.fuse_loss = False)')
        eng.dlog_p.copy_(dlog_p)
        if ctx.needs_input_grad[2]:                   # some flow parameter wants a gradient
            buf, direct = flow._grad_destination()
            if direct:                                # nothing accumulated yet: the graph writes straight into the flow's gradient buffer
                eng.backward_graph(buf).replay()
            else:                                     # gradient accumulation: into the engine's scratch, then add
                eng.backward_graph(eng.dflat).replay()
                buf.add_(eng.dflat)
        else:
            eng.backward_graph(eng.dflat).replay()
        ctx._fin.detach()
        eng.busy, eng.token = False, None
        dfeat = eng.dfeat.clone() if ctx.needs_input_grad[0] else None
        dzd = eng.dz_det.clone() if ctx.needs_input_grad[1] else None
        return dfeat, dzd, None, None, None, None, None, None


def normalize_project(cfg: LossCfg, joints, verts, z, inv_norm: bool, image_size: int = 256):
    """xyz (R,21,3), normalised verts (R,778,3) and uv (R,21,2) for ``MHEnt.sample`` (``network.py:466-514``)."""
    joints = joints.contiguous().float()
    z = z.contiguous().float()
    R = joints.shape[0]
    dev = joints.device
    xyz = torch.empty(R, 21, 3, device=dev)
    uv = torch.empty(R, 21, 2, device=dev)
    verts_n = None
    if verts is not None:
        verts = verts.contiguous().float()
        verts_n = torch.empty_like(verts)
    check(lib().mhe_normalize_project(cfg, ptr(joints), ptr(verts), ptr(z), z.shape[1], R, int(inv_norm), image_size, ptr(xyz),
                                      ptr(verts_n), ptr(uv), stream_ptr(dev)), 'mhe_normalize_project')
    return xyz, verts_n, uv


class MHEntHead(nn.Module):
    """``MHEnt`` downstream of the image feature: flow ``q_z_giv_i``, ``det_head``, ``mano_dec`` and the loss glue.

    ``feat`` (B, 512) is what ``BasicEnc`` produces (``network.py:778-779``); the CNN stays outside.
    """

    def __init__(self, q_z_giv_i_cfg: dict | None = None, mano_data: dict | None = None, image_size: int = 256,
                 feat_dim: int = 512, entropy: bool = True, mano_dir: str = './mano/', synthetic_mano_seed: int | None = None):
        """``mano_data``: MANO constants (``mano_assets.load_mano_pkl`` / ``synthetic_mano``); otherwise ``<mano_dir>/MANO_RIGHT.pkl``
        is read as the reference does (``manolayer.py:61-65``).  A synthetic hand is used only when ``synthetic_mano_seed`` is given."""
        super().__init__()
        cfg = dict(dim=45, tsfm_on=feat_dim, kemb=False, jointN=21, h_dims=[512, 512], num_steps=6)   # CrossModalHand.py:67-69
        if q_z_giv_i_cfg:
            cfg.update(q_z_giv_i_cfg)
        self.q_z_giv_i = RealNVP(**cfg)
        self.mano_dec = ManoLayer(skeidx='RHD', flat_hand_mean=False, ncomps=45, use_pca=True, output_size=image_size,
                                  mask_sz=64, mano_data=mano_data, MANO_dir=mano_dir,
                                  synthetic_seed=synthetic_mano_seed)                                  # network.py:360-363
        self.det_head = nn.Sequential(nn.Linear(feat_dim, feat_dim), nn.ReLU(inplace=True), nn.Linear(feat_dim, 16))  # :380-383
        self.image_size = image_size
        self.entropy = entropy
        self.loss_cfg = default_loss_cfg()
        self.T = 1.0
        self.fuse_loss = True          # get_loss on CUDA tensors replays the engine's captured forward / backward graphs (_FusedLossFn)
        self._fused_pool = {}

    # -- network.py:719-758
    def _sample_q_z_giv_i(self, feat, N=1, temp=1., z0=None, z_det=None, return_log_q=False):
        B = feat.shape[0]
        if z0 is None:
            z0 = torch.randn(N * B, self.q_z_giv_i.dim, device=feat.device) * temp      # == prior.sample * temp (flows.py:339)
        if return_log_q:
            x, log_q = self.q_z_giv_i.sample_with_log_prob(feat, z0, N)
        else:
            x, log_q = self.q_z_giv_i.sample(N * B, temp=temp, logvar=feat, z0=z0), None
        if z_det is None:
            z_det = self.det_head(feat)
        z = _CombineZFn.apply(x, z_det)
        return (z, log_q) if return_log_q else z

    # -- network.py:669-701
    def _reverse_log_q(self, z, feat):
        return self.q_z_giv_i.log_prob(z[:, 3:48].contiguous(), logvar=feat)

    # -- network.py:455-483, 541-558 (+ :497-514)
    def _th_bt_product(self, z, inv_norm=False, want_verts=True):
        dec = self.mano_dec(beta=z[:, 48:58], theta=z[:, :48], want_mesh=want_verts, want_joints=False)
        xyz, verts, uv = normalize_project(self.loss_cfg, dec['mano_joints'], dec['mesh'], z, inv_norm, self.image_size)
        return {'xyz': xyz, 'verts': verts, 'uv': uv}

    # -- network.py:760-831
    def _reverse_kld(self, y: dict, feat, z0=None, z_det=None, N=10, fused=True, want_verts=False):
        """``fused=True`` takes log q of the samples from the sampling pass itself; ``fused=False`` runs the
        reference's second, inverse pass (``network.py:801``).  Both give the same value and gradient."""
        out = {}
        eng = self._fused_engine(feat, N, want_verts) if (fused and self.fuse_loss and self.entropy) else None
        if eng is not None:
            flow = self.q_z_giv_i
            B = feat.shape[0]
            if z0 is None:
                z0 = torch.randn(N * B, flow.dim, device=feat.device)        # == prior.sample * 1.0 (flows.py:339, network.py:720)
            if z_det is None:
                z_det = self.det_head(feat)
            flat = flow._flat_for_autograd(feat.device)
            anchor = flat if flat.requires_grad else feat.new_zeros(())
            log_p, h, qlp, uv, th_norm, bt_norm = _FusedLossFn.apply(feat.float(), z_det.float(), anchor, self, eng, z0.float(),
                                                                    y['crop_uv'].float(), y['vis'].float())
            return {'th_norm': th_norm, 'bt_norm': bt_norm, 'q_log_p_z_giv_y': qlp, 'h_q_z_giv_i': h, 'log_p': log_p, 'uv_mu': uv}
        if fused:
            z, log_q = self._sample_q_z_giv_i(feat, N=N, z0=z0, z_det=z_det, return_log_q=True)
        else:
            z = self._sample_q_z_giv_i(feat, N=N, z0=z0, z_det=z_det)
            log_q = self._reverse_log_q(z, feat)
        out['th_norm'] = z[:, :48].norm(p=2, dim=1)
        out['bt_norm'] = z[:, 48:58].norm(p=2, dim=1)
        dec = self.mano_dec(beta=z[:, 48:58], theta=z[:, :48], want_mesh=want_verts, want_joints=False)
        if not self.entropy:
            log_q = torch.zeros_like(log_q)
        log_p, h, qlp, uv, _ = _ReprojLossFn.apply(dec['mano_joints'], z, log_q, y['crop_uv'], y['vis'], self.loss_cfg)
        out['q_log_p_z_giv_y'] = qlp
        out['h_q_z_giv_i'] = h
        out['log_p'] = log_p
        out['uv_mu'] = uv
        return out

    def _fused_engine(self, feat, N, want_verts):
        """A split-mode engine for this batch shape, or None when the fused node does not apply (CPU tensors, a shape the kernels do not
        cover, both pooled engines still waiting for their backward)."""
        flow = self.q_z_giv_i
        if not feat.is_cuda or not flow._kernel_ok or feat.shape[0] * N > 65536 or feat.shape[0] == 0:
            return None
        from .engine import TrainStep
        key = (feat.shape[0], N, str(feat.device), bool(want_verts), flow.precision, flow._shape.max_split)
        pool = self._fused_pool.setdefault(key, [])
        for eng in pool:
            if not eng.busy:
                return eng
        if len(pool) >= 2:
            return None
        eng = TrainStep(self, feat.shape[0], N, feat.device, want_verts=want_verts, use_graph=False, planes='optimizer', exchange='dense',
                        average_over_ranks=False)
        pool.append(eng)
        return eng

    def _apply(self, fn, *args, **kwargs):
        self._fused_pool = {}            # engines hold device pointers of the parameters: rebuilt after .to() / .cuda()
        return super()._apply(fn, *args, **kwargs)

    def log_prob(self, y: dict, feat, **kw):
        return self._reverse_kld(y, feat, **kw)

    def get_loss(self, feat, y: dict, **kw):
        """``network.py:838-844``; the criterion is ``(-out['log_p']).mean()`` (``criteria.py:55,173``)."""
        return self._reverse_kld(y, feat, **kw)

    # -- network.py:846-883
    @torch.no_grad()
    def sample(self, feat, N=5, temp=0.5, mods=None, y=None, z0=None, z_det=None) -> dict:
        if type(N) == list:
            N, N_quant = N
        else:
            N_quant = N
        B = feat.shape[0]
        if N_quant < N and feat.is_cuda:
            # log q of the samples falls out of the sampling pass itself (log N(z0) - sum s): no second, inverse flow pass
            # (the reference runs one, network.py:866 -> :669; same value, SURVEY.md section 4 identity)
            z, log_q = self._sample_q_z_giv_i(feat, N=N, temp=temp, z0=z0, z_det=z_det, return_log_q=True)
            z, log_q = z.reshape(N, B, 61), log_q.reshape(N, -1)
        else:
            z = self._sample_q_z_giv_i(feat, N=N, temp=temp, z0=z0, z_det=z_det).reshape(N, B, 61)
            log_q = self._reverse_log_q(z.flatten(0, 1), feat).reshape(N, -1) if N_quant < N else None
        if N_quant < N:
            from .metrics import topk_hypotheses
            idx = (topk_hypotheses(log_q, N_quant) if log_q.is_cuda else torch.topk(log_q, N_quant, dim=0)[1])[..., None].repeat(1, 1, 61)
            z = torch.gather(z, 0, idx)
            N = N_quant
        out = {'th_bt': z[..., :58], 'logs_t': z[..., -3:]}
        if mods is None:
            mods = ['xyz', 'uv', 'verts']
        dec = self._th_bt_product(z.reshape(N * B, 61), inv_norm=True, want_verts='verts' in mods)
        for mod in ('verts', 'xyz', 'uv'):
            if mod in mods:
                out[mod] = dec[mod].reshape(N, B, -1)
        if 'verts' in mods:
            out['faces'] = self.mano_dec.mano_faces
        return out
