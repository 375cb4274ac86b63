"""Adam on the flow's flat parameter buffer with the split weight planes refreshed in the same pass.

Mirrors the reference's update, ``torch.optim.Adam(params)`` (``hand/CrossModalHand.py:201``) after
``clip_grad_norm_(encoderRGB.parameters(), 1.)`` (``:462-470``): the clipping norm spans the whole encoder, so the caller passes the
resulting scale (``grad_sqnorm`` gives this module's term).  CUDA only: the kernel is the implementation (``MheError`` otherwise).
"""
from __future__ import annotations

import torch

from . import _lib
from ._lib import check, lib, ptr, stream_ptr


class FlatAdam:
    def __init__(self, flow, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8):
        self.flow, self.lr, self.betas, self.eps = flow, lr, betas, eps
        self.step_count = 0
        self.exp_avg = self.exp_avg_sq = None
        self._sq = None

    def grad_sqnorm(self, dflat: torch.Tensor) -> torch.Tensor:
        """Sum of squares of the flat gradient as a device double (1,): this module's term of the global clipping norm."""
        _lib.require_cuda_f32(dflat)
        if self._sq is None:
            self._sq = torch.zeros(1, dtype=torch.float64, device=dflat.device)
        check(lib().mhe_flow_grad_sqnorm(self.flow._shape, ptr(dflat), ptr(self._sq), stream_ptr(dflat.device)), 'mhe_flow_grad_sqnorm')
        return self._sq

    def step(self, dflat: torch.Tensor, grad_scale: float = 1.0, refresh_planes: bool = True):
        """One update of every flow parameter from the flat gradient (layout of ``mhe_flow_param_offset``)."""
        dev = dflat.device
        _lib.require_cuda_f32(dflat)
        flat = self.flow.flat_parameters(dev)
        if self.exp_avg is None:
            self.exp_avg, self.exp_avg_sq = torch.zeros_like(flat), torch.zeros_like(flat)
        packed = self.flow.packed_weights(dev) if refresh_planes and self.flow.precision != 'fp32' else None
        self.step_count += 1
        check(lib().mhe_flow_adam_step(self.flow._shape, ptr(flat), ptr(dflat.contiguous()), ptr(self.exp_avg), ptr(self.exp_avg_sq),
                                       ptr(packed) if packed is not None else None, self.step_count, self.lr, self.betas[0], self.betas[1],
                                       self.eps, grad_scale, stream_ptr(dev)), 'mhe_flow_adam_step')
        # the update went through raw pointers (no version counter moved): with the planes refreshed here the flow's cached set IS
        # current - record that; without, force a re-pack before the next tensor-core pass
        self.flow.mark_parameters_changed()
        if packed is not None:
            self.flow._packed_sig = self.flow._packed_signature(flat)
