"""Drop-in ``RealNVP`` — the reference's conditional flow (``hand/flows.py:125-362``) on the B200 kernels.

Same constructor, methods, attributes and state-dict keys as the reference class, so
``MHEnt`` (``hand/network.py:341,692,733``) and released checkpoints (``ent_ho3d.pth``) keep working:
``t.{i}.l.{j}.*``, ``t.{i}.c.{j}.*``, ``s.…``, ``mask``.  The parameters are views into one flat fp32
buffer (layout in ``include/mhentropy_b200.h``) that the CUDA kernels read directly; gradients come
back in the same flat layout.

The kernel path covers the configuration the reference ships (int ``tsfm_on`` = conditional flow,
no ``kemb`` / partitioner, equal hidden widths, ``weights=None``).  Other constructor options keep
their reference semantics through stock PyTorch ops (API coverage, not a fallback of the measured
path).  CUDA tensors always take the kernel path and fail loudly if the library is missing.
"""
from __future__ import annotations

import math
import os

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F
from torch import distributions

from . import _lib
from ._lib import FlowShape, check, lib, ptr, stream_ptr

LEAKY_SLOPE = 0.01


class _nets(nn.Module):
    """Coupling MLP container (reference ``flows.py:75-122``): holds ``l.{0,1,2}`` and ``c.{0,1}``.

    ``forward`` is the stock-PyTorch evaluation used only off the kernel path (CPU tensors or
    unsupported configurations).
    """

    def __init__(self, dim, cond_dim=0, h_dims=(64, 64), s=True):
        super().__init__()
        self.cond_dim = cond_dim
        self.s = s
        self.l = nn.ModuleList([nn.Linear(dim, h_dims[0]), nn.Linear(h_dims[0], h_dims[1]), nn.Linear(h_dims[1], dim)])
        if cond_dim:
            self.c = nn.ModuleList([nn.Linear(cond_dim, h) for h in h_dims])

    def forward(self, x, cond=None):
        h = self.l[0](x)
        for i in range(2):
            if self.cond_dim > 0:
                h = h + self.c[i](cond)
            h = self.l[i + 1](F.leaky_relu(h, LEAKY_SLOPE))
        return torch.tanh(h) if self.s else h


_WHICH = {('l', 0, 'weight'): 0, ('l', 0, 'bias'): 1, ('l', 1, 'weight'): 2, ('l', 1, 'bias'): 3,
          ('l', 2, 'weight'): 4, ('l', 2, 'bias'): 5, ('c', 0, 'weight'): 6, ('c', 0, 'bias'): 7,
          ('c', 1, 'weight'): 8, ('c', 1, 'bias'): 9}


class _FlatAnchor(torch.autograd.Function):
    """Autograd bridge between the 240 reference-named parameters and the flat buffer the kernels read.

    The flat buffer enters the graph through ONE scalar anchor that requires grad, so the kernels' backward functions run
    whenever a parameter wants a gradient; they write straight into the flow's persistent flat gradient buffer, of which
    every parameter's ``.grad`` is a view (``RealNVP.grad_buffer``).  Nothing flows back through this node: routing 240
    tensors through autograd cost more host time per step than the kernels themselves.
    """

    @staticmethod
    def forward(ctx, anchor, flow):
        return flow._flat.detach()

    @staticmethod
    def backward(ctx, gflat):
        return None, None


class _CondFn(torch.autograd.Function):
    """cp = hoisted conditioning projections of every layer (``flows.py:107-109``)."""

    @staticmethod
    def forward(ctx, feat, flat, shape, packed, flow=None):
        feat = feat.contiguous()
        _lib.require_cuda_f32(feat, flat)
        B = feat.shape[0]
        dev = feat.device
        cp = torch.empty(B, lib().mhe_flow_cp_floats_per_image(shape), device=dev, dtype=torch.float32)
        wsb = lib().mhe_flow_cond_workspace_bytes(shape, B) if packed is not None else 0
        ws = _lib.WORKSPACE.get(wsb, dev, 'cond') if wsb else None
        check(lib().mhe_flow_cond_fwd(shape, ptr(flat), ptr(packed), ptr(feat), B, ptr(cp), ptr(ws), wsb, stream_ptr(dev)), 'mhe_flow_cond_fwd')
        ctx.save_for_backward(feat, flat)
        ctx.shape, ctx.packed, ctx.flow = shape, packed, flow
        return cp

    @staticmethod
    def backward(ctx, dcp):
        feat, flat = ctx.saved_tensors
        dcp = dcp.contiguous()
        dev = feat.device
        # parameter gradients go straight into the flow's persistent flat gradient buffer (the parameters' .grad are views of it)
        dflat, store = ctx.flow._grad_target('cond') if ctx.flow is not None else (torch.zeros_like(flat), True)
        dfeat = torch.empty_like(feat) if ctx.needs_input_grad[0] else None
        B = feat.shape[0]
        wsb = lib().mhe_flow_cond_workspace_bytes(ctx.shape, B) if ctx.packed is not None else 0
        ws = _lib.WORKSPACE.get(wsb, dev, 'cond') if wsb else None
        check(lib().mhe_flow_set_async(2 if store else 0), 'mhe_flow_set_async')   # bit 1: the weight slots are zero -> store, no read
        try:
            check(lib().mhe_flow_cond_bwd(ctx.shape, ptr(flat), ptr(ctx.packed), ptr(feat), ptr(dcp), B, ptr(dflat), ptr(dfeat), ptr(ws), wsb,
                                          stream_ptr(dev)), 'mhe_flow_cond_bwd')
        finally:
            check(lib().mhe_flow_set_async(0), 'mhe_flow_set_async')
        return dfeat, (dflat if ctx.flow is None else None), None, None, None


class _FlowPassFn(torch.autograd.Function):
    """One pass through the coupling layers; returns (out, logdet)."""

    @staticmethod
    def forward(ctx, inp, cp, flat, mask, shape, direction, B, packed, flow=None):
        inp = inp.contiguous()
        _lib.require_cuda_f32(inp, cp, flat, mask)
        R, D = inp.shape
        dev = inp.device
        out = torch.empty_like(inp)
        logdet = torch.empty(R, device=dev, dtype=torch.float32)
        need_grad = any(ctx.needs_input_grad[:3])
        tcore = int(packed is not None)
        saved = torch.empty(lib().mhe_flow_saved_bytes(shape, R, tcore), device=dev, dtype=torch.uint8) if need_grad else None
        wsb = lib().mhe_flow_workspace_bytes(shape, R, int(packed is not None))
        ws = _lib.WORKSPACE.get(wsb, dev)
        check(lib().mhe_flow_pass_fwd(shape, ptr(flat), ptr(packed), ptr(mask), ptr(cp), ptr(inp), R, B, direction, ptr(out), ptr(logdet),
                                      ptr(saved), ptr(ws), wsb, stream_ptr(dev)), 'mhe_flow_pass_fwd')
        if need_grad:
            ctx.save_for_backward(cp, flat, mask, saved)
        ctx.shape, ctx.direction, ctx.B, ctx.packed, ctx.dims, ctx.flow = shape, direction, B, packed, (R, D), flow
        return out, logdet

    @staticmethod
    def backward(ctx, dout, dlogdet):
        cp, flat, mask, saved = ctx.saved_tensors
        shape = ctx.shape
        R, D = ctx.dims
        dev = saved.device
        dout = torch.zeros(R, D, device=dev) if dout is None else dout.contiguous()
        dlogdet = None if dlogdet is None else dlogdet.contiguous()
        din = torch.empty(R, D, device=dev, dtype=torch.float32)
        dflat, store = ctx.flow._grad_target('pass') if ctx.flow is not None else (torch.zeros_like(flat), True)
        dcp = torch.zeros_like(cp)
        wsb = lib().mhe_flow_workspace_bytes(shape, R, int(ctx.packed is not None))
        ws = _lib.WORKSPACE.get(wsb, dev)
        check(lib().mhe_flow_set_async(2 if store else 0), 'mhe_flow_set_async')
        try:
            check(lib().mhe_flow_pass_bwd(shape, ptr(flat), ptr(ctx.packed), ptr(mask), ptr(cp), ptr(saved), R, ctx.B, ctx.direction, ptr(dout),
                                          ptr(dlogdet), 1.0, ptr(din), ptr(dflat), ptr(dcp), ptr(ws), wsb, stream_ptr(dev)),
                  'mhe_flow_pass_bwd')
        finally:
            check(lib().mhe_flow_set_async(0), 'mhe_flow_set_async')
        return din, dcp, (dflat if ctx.flow is None else None), None, None, None, None, None, None


class _StdNormalLogpFn(torch.autograd.Function):
    """log N(z; 0, I) + logdet (``flows.py:320``)."""

    @staticmethod
    def forward(ctx, z, logdet):
        z = z.contiguous()
        logdet = logdet.contiguous()
        _lib.require_cuda_f32(z, logdet)
        R, D = z.shape
        logp = torch.empty(R, device=z.device, dtype=torch.float32)
        check(lib().mhe_std_normal_logp_fwd(ptr(z), ptr(logdet), 1.0, R, D, ptr(logp), stream_ptr(z.device)), 'mhe_std_normal_logp_fwd')
        ctx.save_for_backward(z)
        return logp

    @staticmethod
    def backward(ctx, dlogp):
        (z,) = ctx.saved_tensors
        dlogp = dlogp.contiguous()
        dz = torch.empty_like(z)
        check(lib().mhe_std_normal_logp_bwd(ptr(z), ptr(dlogp), z.shape[0], z.shape[1], ptr(dz), stream_ptr(z.device)),
              'mhe_std_normal_logp_bwd')
        return dz, dlogp


def lib_has_tc(shape) -> bool:
    try:
        return lib().mhe_flow_packed_bytes(shape) > 0
    except _lib.MheError:
        return False


class RealNVP(nn.Module):
    """Conditional RealNVP, API-compatible with reference ``flows.RealNVP`` (``flows.py:125-362``)."""

    def __init__(self, nets=_nets, nett=_nets, mask=None, prior=None, dim=63, tsfm_on=None, kemb=False, jointN=21,
                 h_dims=(64, 64), num_steps=3, cond_mapping_dims=None):
        super().__init__()
        if dim == 1:
            raise ValueError
        if kemb:
            raise NotImplementedError('kemb (joint-independent embeddings) is outside the accelerated path')
        self.dim = dim
        self.jointN = jointN
        if mask is None:
            a = [0] * (dim // 2) + [1] * (dim - dim // 2)
            b = [1 - v for v in a]
            mask = torch.from_numpy(np.array([a, b] * num_steps).astype(np.float32))   # flows.py:152-155
        if prior is None:
            prior = distributions.MultivariateNormal(torch.zeros(dim), torch.eye(dim))   # flows.py:156-157
        self.tsfm_on = tsfm_on
        cond_dim = tsfm_on if type(tsfm_on) == int else 0
        partitioner = nn.ModuleList()
        for in_f, out_f in (cond_mapping_dims or []):
            assert out_f % jointN == 0
            partitioner.append(nn.Linear(in_f, out_f))
        self.joint_feat_partitioner = partitioner
        self.prior = prior
        self.register_buffer('mask', mask)
        h_dims = list(h_dims)
        self.t = nn.ModuleList([nett(dim, cond_dim=cond_dim, h_dims=h_dims, s=False) for _ in range(len(mask))])
        self.s = nn.ModuleList([nets(dim, cond_dim=cond_dim, h_dims=h_dims) for _ in range(len(mask))])
        self.scale = 1.
        self.h_dims = h_dims
        self.cond_dim = cond_dim
        self._structure_ok = (cond_dim > 0 and len(h_dims) == 2 and h_dims[0] == h_dims[1] and 2 <= dim <= 64
                              and len(partitioner) == 0 and nets is _nets and nett is _nets)
        self._flat = None
        self._slots = None
        self._grad_flat = None
        self._grad_views = None
        self._grad_writes = {'pass': 0, 'cond': 0}
        self._fwd_gen = 0
        self._grad_zero_gen = -1
        self._anchor = None
        self._packed = None
        self._packed_sig = None
        self._param_epoch = 0
        self._refresh_mask_info()
        # 'bf16x3': tcgen05 tensor cores, split-bf16 (hi*hi + hi*lo + lo*hi, fp32 accumulate); 'fp32': CUDA cores, exact
        self.precision = 'bf16x3' if (self._kernel_ok and lib_has_tc(self._shape)) else 'fp32'

    # ------------------------------------------------------------------ coupling masks
    def _refresh_mask_info(self):
        """What the kernels need to know about the masks (``flows.py:152-155``; a user may pass any ``mask``, ``:131``).

        The kernels implement the coupling for {0,1} masks only (the reference multiplies by the float mask, so a fractional
        mask is a different function): anything else is outside the accelerated path.  ``max_split`` - the largest number of
        transformed or conditioning dims of a layer - travels in the shape: the cluster-fused kernels exchange at most 24 dims per
        layer and side, larger splits (e.g. 30/15) take the per-GEMM tensor-core path.
        """
        m = self.mask.detach().float().cpu()
        binary = bool(((m == 0) | (m == 1)).all())
        ones = m.sum(1)
        max_split = int(max(ones.max().item(), (m.shape[1] - ones).max().item())) if binary and m.numel() else 0
        self._mask_binary = binary
        self._kernel_ok = self._structure_ok and binary and m.shape[1] == self.dim
        self._shape = FlowShape(self.dim, self.h_dims[0], self.cond_dim, m.shape[0], max_split) if self._kernel_ok else None

    def _load_from_state_dict(self, *args, **kwargs):
        super()._load_from_state_dict(*args, **kwargs)
        self._refresh_mask_info()           # a checkpoint may carry another mask
        self.mark_parameters_changed()

    def mark_parameters_changed(self):
        """Tell the flow that its parameters were modified behind autograd's back (raw-pointer optimizer step, ``.data`` edits): the
        split weight planes are re-packed before the next tensor-core pass.  ``FlatAdam.step`` calls it; in-place updates through
        PyTorch bump the tensors' version counters and are noticed without it."""
        self._param_epoch += 1

    # ------------------------------------------------------------------ flat parameter storage
    def _named_flow_params(self):
        for net_id, name in ((0, 's'), (1, 't')):
            for i, net in enumerate(getattr(self, name)):
                for kind in ('l', 'c'):
                    for j, lin in enumerate(getattr(net, kind)):
                        for pname in ('weight', 'bias'):
                            yield (i, net_id, _WHICH[(kind, j, pname)]), getattr(lin, pname)

    def _adopt(self, device):
        """Move every parameter into one flat buffer on ``device``; parameters become views of it."""
        shape = self._shape
        total = lib().mhe_flow_param_floats(shape)
        flat = torch.zeros(total, device=device, dtype=torch.float32)
        slots = []
        for (i, n, which), p in self._named_flow_params():
            off = lib().mhe_flow_param_offset(shape, i, n, which)
            view = flat[off:off + p.numel()].view(p.shape)
            view.copy_(p.data)
            p.data = view
            slots.append((p, off, p.numel(), tuple(p.shape)))
        self._flat, self._slots = flat, slots
        self._grad_flat = self._grad_views = None
        self._anchor = torch.zeros((), device=device, requires_grad=True)

    def _flat_is_current(self, device) -> bool:
        if self._flat is None or self._flat.device != device:
            return False
        base = self._flat.data_ptr()
        for p, off, _, _ in (self._slots[0], self._slots[len(self._slots) // 2], self._slots[-1]):
            if p.data_ptr() != base + 4 * off:
                return False
        return True

    def flat_parameters(self, device=None) -> torch.Tensor:
        """The flat fp32 parameter buffer the kernels read (adopting the parameters if needed)."""
        device = torch.device(device) if device is not None else self.mask.device
        if device.type == 'cuda' and device.index is None:      # 'cuda' == the current device: tensors always report an index
            device = torch.device('cuda', torch.cuda.current_device())
        if not self._flat_is_current(device):
            self._adopt(device)
        return self._flat

    def packed_weights(self, device=None):
        """Split-bf16 planes of the weights for the tensor-core path (None on the fp32 path); refreshed
        whenever the parameters have been modified in place (optimizer step, load_state_dict)."""
        if self.precision == 'fp32':
            return None
        if self.precision != 'bf16x3':
            raise ValueError(f"precision must be 'fp32' or 'bf16x3', got {self.precision!r}")
        flat = self.flat_parameters(device)
        sig = self._packed_signature(flat)
        if self._packed is None or self._packed.device != flat.device or sig != self._packed_sig:
            nbytes = lib().mhe_flow_packed_bytes(self._shape)
            if nbytes == 0:
                raise _lib.MheError('this flow shape is outside the tensor-core path; set precision="fp32"')
            if self._packed is None or self._packed.device != flat.device:
                self._packed = torch.empty(nbytes, dtype=torch.uint8, device=flat.device)
            check(lib().mhe_flow_pack_weights(self._shape, ptr(flat), ptr(self._packed), 3, stream_ptr(flat.device)), 'mhe_flow_pack_weights')
            self._packed_sig = sig
        return self._packed

    def _packed_signature(self, flat):
        # every parameter's version counter (in-place updates through PyTorch) + the explicit epoch (raw-pointer updates)
        return (flat.data_ptr(), flat._version, self._param_epoch, sum(p._version for p, _, _, _ in self._slots))

    def _split_flat(self, gflat):
        return [gflat[off:off + n].view(shape) for _, off, n, shape in self._slots]

    def grad_buffer(self, device=None) -> torch.Tensor:
        """The persistent flat gradient buffer (layout of ``mhe_flow_param_offset``).  After ``backward()`` every parameter's
        ``.grad`` is a view into it, so a data-parallel exchange is ONE all-reduce of this tensor and ``FlatAdam.step`` consumes it
        directly."""
        flat = self.flat_parameters(device)
        if self._grad_flat is None or self._grad_flat.device != flat.device:
            self._grad_flat = torch.zeros_like(flat)
            self._grad_views = self._split_flat(self._grad_flat)
            self._grad_writes = {'pass': 0, 'cond': 0}
        return self._grad_flat

    def _grads_attached(self) -> bool:
        slots, views = self._slots, self._grad_views
        if views is None:
            return False
        for i in (0, len(slots) // 2, len(slots) - 1):      # zero_grad(set_to_none=True) / a foreign .grad detaches all or nothing in practice
            p = slots[i][0]
            if p.requires_grad and p.grad is not views[i]:
                return False
        return True

    def _grad_target(self, family: str):
        """Where a backward function accumulates the parameter gradients, and whether it may STORE the weight slots of its family
        (``'pass'``: W0/W1/W2, ``'cond'``: Cw) instead of read-modify-write.

        ``.grad is None`` (``zero_grad(set_to_none=True)``, the first step) re-attaches the views and zeroes the buffer with one
        memset; the first writer of each family in that same backward sweep may then store.  Attached views mean the caller
        accumulates (or has zeroed them in place), so the kernels add."""
        buf = self.grad_buffer()
        if not self._grads_attached():
            buf.zero_()
            for (p, _, _, _), v in zip(self._slots, self._grad_views):
                if p.requires_grad:
                    if p.grad is not None and p.grad is not v:
                        v.add_(p.grad)                   # a gradient accumulated elsewhere is carried over
                    p.grad = v
            self._grad_writes = {'pass': 0, 'cond': 0}
            self._grad_zero_gen = self._fwd_gen
        store = self._grad_zero_gen == self._fwd_gen and self._grad_writes[family] == 0
        self._grad_writes[family] += 1
        return buf, store

    def _grad_destination(self):
        """For a backward that OVERWRITES its destination (the engine's captured backward graph): ``(buffer, direct)``.
        ``direct``: no gradient is attached yet (first step / ``zero_grad(set_to_none=True)``) - the views are attached and the graph may
        write straight into the flow's gradient buffer.  Otherwise the caller must write to scratch and ``add_`` (gradient accumulation)."""
        buf = self.grad_buffer()
        if self._grads_attached():
            return buf, False
        carry = []
        for (p, _, _, _), v in zip(self._slots, self._grad_views):      # one pass: attach the views, note gradients held elsewhere
            if p.requires_grad:
                g = p.grad
                if g is not None and g is not v:
                    carry.append((v, g))
                p.grad = v
        if carry:                                   # gradients accumulated elsewhere are carried over into the flat buffer
            buf.zero_()
            for v, g in carry:
                v.copy_(g)
        self._grad_writes = {'pass': 1, 'cond': 1}
        return buf, not carry

    def _flat_for_autograd(self, device):
        flat = self.flat_parameters(device)
        if torch.is_grad_enabled() and any(p.requires_grad for p, _, _, _ in self._slots):
            self._fwd_gen += 1                       # a new forward: gradients written after it belong to a new backward sweep
            return _FlatAnchor.apply(self._anchor, self)
        return flat

    def _use_kernels(self, x) -> bool:
        if x.is_cuda:
            if not self._kernel_ok:
                raise _lib.MheError('this RealNVP configuration is outside the accelerated path'
                                    + ('' if self._mask_binary else ' (the coupling mask is not exactly {0,1})')
                                    + '; run it on CPU tensors through the stock implementation or use the shipped configuration')
            return True
        return False

    # ------------------------------------------------------------------ reference API
    def cond_projections(self, cond):
        """Hoisted ``c.{0,1}(cond)`` of all layers, (B, L*4*H); kernel path only."""
        flat = self._flat_for_autograd(cond.device)
        return _CondFn.apply(cond.float(), flat, self._shape, self.packed_weights(cond.device), self)

    def _rowcond_applies(self, inp, cond) -> bool:
        """One conditioning row per flow row, nothing to differentiate, long batch on the tensor-core path: the pass that contracts the
        conditioning projections inside the coupling GEMMs (``mhe_flow_pass_fwd_rowcond``; ``MHE_FLOW_ROWCOND=0`` disables it)."""
        if cond is None or cond.shape[0] != inp.shape[0] or inp.shape[0] == 0 or os.environ.get('MHE_FLOW_ROWCOND') == '0':
            return False
        if torch.is_grad_enabled() and (inp.requires_grad or cond.requires_grad or any(p.requires_grad for p, _, _, _ in self._slots)):
            return False
        return self.packed_weights(inp.device) is not None and bool(lib().mhe_flow_rowcond_supported(self._shape, inp.shape[0]))

    def _pass_rowcond(self, inp, cond, direction):
        dev = inp.device
        flat, packed = self.flat_parameters(dev), self.packed_weights(dev)
        inp, feat = inp.float().contiguous(), cond.float().contiguous()
        _lib.require_cuda_f32(inp, feat, flat, self.mask)
        R = inp.shape[0]
        out = torch.empty_like(inp)
        logdet = torch.empty(R, device=dev, dtype=torch.float32)
        wsb = lib().mhe_flow_rowcond_workspace_bytes(self._shape, R)
        ws = _lib.WORKSPACE.get(wsb, dev)
        check(lib().mhe_flow_pass_fwd_rowcond(self._shape, ptr(flat), ptr(packed), ptr(self.mask), ptr(feat), ptr(inp), R, direction, ptr(out),
                                              ptr(logdet), ptr(ws), wsb, stream_ptr(dev)), 'mhe_flow_pass_fwd_rowcond')
        return out, logdet

    def _pass(self, inp, cond, direction, cp=None, images=None):
        if cp is None and self._rowcond_applies(inp, cond):
            return self._pass_rowcond(inp, cond, direction)
        flat = self._flat_for_autograd(inp.device)
        packed = self.packed_weights(inp.device)
        if cp is None:
            cp = _CondFn.apply(cond.float(), flat, self._shape, packed, self)
            images = cond.shape[0]
        return _FlowPassFn.apply(inp.float(), cp, flat, self.mask, self._shape, direction, images, packed, self)

    def forward_p(self, z, cond=None):
        """z -> x (``flows.py:210-217``)."""
        if self._use_kernels(z):
            return self._pass(z, cond, 0)[0]
        x = z
        for i in range(len(self.t)):
            x_ = x * self.mask[i]
            s = self.s[i](x_, cond=cond) * (1 - self.mask[i])
            t = self.t[i](x_, cond=cond) * (1 - self.mask[i])
            x = x_ + (1 - self.mask[i]) * (x * torch.exp(s) + t)
        return x

    def backward_p(self, x, cond=None):
        """x -> (z, log|det|) (``flows.py:219-227``)."""
        if self._use_kernels(x):
            return self._pass(x, cond, 1)
        log_det, z = x.new_zeros(x.shape[0]), x
        for i in reversed(range(len(self.t))):
            z_ = self.mask[i] * z
            s = self.s[i](z_, cond=cond) * (1 - self.mask[i])
            t = self.t[i](z_, cond=cond) * (1 - self.mask[i])
            z = (1 - self.mask[i]) * (z - t) * torch.exp(-s) + z_
            log_det = log_det - s.sum(dim=1)
        return z, log_det

    def make_cond(self, feat):
        """``flows.py:229-269`` for the supported case (no kemb, no partitioner): a reshape."""
        bs = feat.shape[0]
        bs1 = bs * self.jointN if self.dim in [2, 3] else bs
        if self.dim <= 3 and len(self.joint_feat_partitioner):
            raise NotImplementedError('joint feature partitioner is outside the accelerated path')
        return feat.reshape(bs1, -1)

    def log_prob(self, x, mu=None, logvar=None, return_dict=False, weights=None, return_z=False):
        """``flows.py:271-331``.  ``logvar`` carries the features when ``tsfm_on`` is an int.

        Extension: ``logvar`` may hold one feature row per image, (B, F), while ``x`` holds R = S*B
        hypothesis-major rows (row r belongs to image r % B); the conditioning is then projected once
        per image instead of once per row.  With R rows of features (the reference's
        ``feat.repeat(N, 1)``) the behaviour is the reference's.
        """
        if type(self.tsfm_on) != int:
            raise NotImplementedError("tsfm_on in {'x','z',None} is outside the accelerated path")
        if weights is not None and bool((1 - weights.float()).count_nonzero()):
            raise NotImplementedError       # flows.py:284-285
        bs = x.shape[0]
        x = x.reshape(-1, self.dim) / self.scale
        cond = self.make_cond(logvar)
        if self._use_kernels(x):
            z, logdet = self._pass(x, cond, 1)
            loss = _StdNormalLogpFn.apply(z, logdet).view(bs, -1).sum(1)
        else:
            z, logdet = self.backward_p(x, cond=cond)
            loss = (-0.5 * (z * z).sum(-1) - 0.5 * self.dim * math.log(2 * math.pi) + logdet).view(bs, -1).sum(1)
        if return_z:
            return z, loss
        if return_dict:
            return {'loss': loss}
        return loss

    def sample(self, batchSize, temp=0.7, mu=None, logvar=None, return_z=False, z0=None):
        """``flows.py:333-359``.  ``z0`` (extension) supplies the prior draw instead of sampling it here."""
        if type(self.tsfm_on) != int:
            raise NotImplementedError("tsfm_on in {'x','z',None} is outside the accelerated path")
        if z0 is None:
            z0 = self.prior.sample((batchSize,)).to(logvar.device) * temp     # flows.py:339
        z = z0
        cond = self.make_cond(logvar)
        x = self.forward_p(z, cond=cond) * self.scale
        bs = logvar.shape[0] if self.dim in [2, 3] else z0.shape[0]
        if return_z:
            return x.view(bs, -1), z0.view(bs, -1)
        return x.view(bs, -1)

    def sample_with_log_prob(self, feat, z0, S):
        """Fused extension: x = flow(z0 | feat) and log q(x | feat) from ONE pass (SURVEY.md §4 identity).

        feat (B, F) one row per image; z0 (S*B, dim) hypothesis-major.  Returns x (S*B, dim), log_q (S*B,).
        """
        flat = self._flat_for_autograd(z0.device)
        packed = self.packed_weights(z0.device)
        cp = _CondFn.apply(feat.float(), flat, self._shape, packed, self)
        x, logdet = _FlowPassFn.apply(z0.float(), cp, flat, self.mask, self._shape, 0, feat.shape[0], packed, self)
        log_q = _StdNormalLogpFn.apply(z0.float(), -logdet)
        return x * self.scale, log_q

    def forward(self, x):
        return self.log_prob(x)
