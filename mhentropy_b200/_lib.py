"""ctypes binding of ``libmhentropy_b200.so`` (the C ABI declared in ``include/mhentropy_b200.h``).

There is no CPU fallback: if the library is missing or a call fails, this raises.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_float, c_int, c_longlong, c_size_t, c_void_p

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libmhentropy_b200.so')


class FlowShape(Structure):
    _fields_ = [('dim', c_int), ('hidden', c_int), ('cond', c_int), ('layers', c_int), ('max_split', c_int)]


class ManoConsts(Structure):
    _fields_ = [(n, c_void_p) for n in
                ('comps', 'hands_mean', 'v_template', 'shapedirs', 'posedirs_t', 'jreg', 'weights', 'jt', 'js', 'pose_tables',
                 'posedirs_planes')]


class LossCfg(Structure):
    _fields_ = [('laplace_b', c_float), ('th45_box', c_float), ('th45_alpha', c_float), ('th3_radius', c_float),
                ('th3_alpha', c_float), ('bt_box', c_float), ('bt_alpha', c_float), ('root_idx', c_int), ('norm_idx', c_int)]


class MheError(RuntimeError):
    pass


_P = c_void_p
_SIGNATURES = {
    'mhe_last_error_string': (c_char_p, []),
    'mhe_version': (c_int, []),
    'mhe_built_for_sm': (c_int, []),
    'mhe_kernel_launch_count': (c_longlong, []),
    'mhe_probe_configure': (c_int, [c_char_p, c_int]),
    'mhe_probe_reset': (c_int, []),
    'mhe_probe_read': (c_int, [POINTER(c_float), POINTER(c_int)]),
    'mhe_flow_param_floats': (c_size_t, [FlowShape]),
    'mhe_flow_param_offset': (c_size_t, [FlowShape, c_int, c_int, c_int]),
    'mhe_flow_cp_floats_per_image': (c_size_t, [FlowShape]),
    'mhe_flow_workspace_bytes': (c_size_t, [FlowShape, c_int, c_int]),
    'mhe_flow_saved_bytes': (c_size_t, [FlowShape, c_int, c_int]),
    'mhe_flow_cond_workspace_bytes': (c_size_t, [FlowShape, c_int]),
    'mhe_flow_packed_bytes': (c_size_t, [FlowShape]),
    'mhe_flow_pack_weights': (c_int, [FlowShape, _P, _P, c_int, _P]),
    'mhe_flow_cond_fwd_uses_planes': (c_int, [FlowShape, c_int]),
    'mhe_flow_zero_bias_grads': (c_int, [FlowShape, _P, _P]),
    'mhe_flow_pass_bwd_prepare': (c_int, [FlowShape, _P, _P, c_int, c_int, _P, c_size_t, _P]),
    'mhe_flow_pass_cond_bwd': (c_int, [FlowShape, _P, _P, _P, _P, _P, c_int, c_int, c_int, _P, _P, c_float, _P, _P, _P, _P, _P, _P, c_size_t, _P, c_size_t, _P]),
    'mhe_flow_cond_wgrad': (c_int, [FlowShape, _P, _P, c_int, _P, _P, c_size_t, _P]),
    'mhe_flow_grad_sqnorm': (c_int, [FlowShape, _P, _P, _P]),
    'mhe_flow_rowcond_supported': (c_int, [FlowShape, c_int]),
    'mhe_flow_rowcond_workspace_bytes': (c_size_t, [FlowShape, c_int]),
    'mhe_flow_pass_fwd_rowcond': (c_int, [FlowShape, _P, _P, _P, _P, _P, c_int, c_int, _P, _P, _P, c_size_t, _P]),
    'mhe_sum_shards': (c_int, [_P, _P, c_int, c_int, c_size_t, c_size_t, _P]),
    'mhe_flow_adam_step': (c_int, [FlowShape, _P, _P, _P, _P, _P, c_int, ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.c_double,
                                   ctypes.c_double, _P]),
    'mhe_flow_pass_is_fused': (c_int, [FlowShape, c_int]),
    'mhe_flow_bwd_chunk_count': (c_int, [FlowShape, c_int]),
    'mhe_flow_bwd_chunk_layers': (c_int, [FlowShape, c_int, c_int, c_int, POINTER(c_int), POINTER(c_int)]),
    'mhe_flow_join_chunk': (c_int, [_P, c_int]),
    'mhe_flow_cond_fwd': (c_int, [FlowShape, _P, _P, _P, c_int, _P, _P, c_size_t, _P]),
    'mhe_flow_cond_bwd': (c_int, [FlowShape, _P, _P, _P, _P, c_int, _P, _P, _P, c_size_t, _P]),
    'mhe_flow_pass_fwd': (c_int, [FlowShape, _P, _P, _P, _P, _P, c_int, c_int, c_int, _P, _P, _P, _P, c_size_t, _P]),
    'mhe_flow_pass_bwd': (c_int, [FlowShape, _P, _P, _P, _P, _P, c_int, c_int, c_int, _P, _P, c_float, _P, _P, _P, _P, c_size_t, _P]),
    'mhe_flow_set_async': (c_int, [c_int]),
    'mhe_flow_join': (c_int, [_P]),
    'mhe_std_normal_logp_fwd': (c_int, [_P, _P, c_float, c_int, c_int, _P, _P]),
    'mhe_std_normal_logp_bwd': (c_int, [_P, _P, c_int, c_int, _P, _P]),
    'mhe_mano_pose_tables_floats': (c_size_t, []),
    'mhe_mano_pack_pose_tables': (c_int, [POINTER(ManoConsts), _P, _P]),
    'mhe_mano_posedirs_planes_bytes': (c_size_t, []),
    'mhe_mano_pack_posedirs_planes': (c_int, [POINTER(ManoConsts), _P, _P]),
    'mhe_mano_workspace_bytes': (c_size_t, [c_int, c_int]),
    'mhe_mano_fwd': (c_int, [POINTER(ManoConsts), _P, c_int, _P, c_int, c_int, c_int, _P, _P, _P, _P, c_size_t, _P]),
    'mhe_mano_bwd': (c_int, [POINTER(ManoConsts), _P, c_int, _P, c_int, c_int, c_int, _P, _P, _P, _P, c_int, _P, c_int,
                             c_int, _P, c_size_t, _P]),
    'mhe_combine_z_fwd': (c_int, [_P, _P, c_int, c_int, _P, _P]),
    'mhe_combine_z_bwd': (c_int, [_P, c_int, c_int, _P, _P, _P]),
    'mhe_reproj_loss_fwd': (c_int, [POINTER(LossCfg), _P, _P, _P, _P, _P, c_int, c_int, _P, _P, _P, _P, _P, _P, _P]),
    'mhe_reproj_loss_bwd': (c_int, [POINTER(LossCfg), _P, _P, _P, _P, c_int, c_int, _P, _P, _P, _P, _P, _P]),
    'mhe_image_loss_reduce': (c_int, [_P, _P, c_int, c_int, _P, _P, _P, _P, _P]),
    'mhe_hypothesis_rows_fwd_bwd': (c_int, [POINTER(ManoConsts), POINTER(LossCfg), _P, _P, _P, _P, _P, c_int, c_int, c_int, c_float, _P, _P, _P, _P, _P,
                                            _P, _P, _P]),
    'mhe_hypothesis_metrics_workspace_bytes': (c_size_t, [c_int]),
    'mhe_hypothesis_metrics': (c_int, [_P, _P, _P, _P, _P, _P, c_int, c_int, c_int, c_float, _P, _P, c_size_t, _P]),
    'mhe_topk_hypotheses': (c_int, [_P, c_int, c_int, c_int, _P, _P]),
    'mhe_split_planes': (c_int, [_P, c_int, c_int, _P, c_int, c_int, c_int, c_int, c_int, _P]),
    'mhe_tc_gemm_raw': (c_int, [_P, _P, _P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, _P]),
    'mhe_normalize_project': (c_int, [POINTER(LossCfg), _P, _P, _P, c_int, c_int, c_int, c_int, _P, _P, _P, _P]),
}

_lib = None


def exported_symbols() -> list[str]:
    """Every entry point ``include/mhentropy_b200.h`` declares."""
    return sorted(_SIGNATURES)


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise MheError(
                f'{LIB_PATH} is missing: build it with `python -m mhentropy_b200.build` '
                '(there is no CPU or PyTorch fallback for this path)')
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(status: int, what: str) -> None:
    if status != 0:
        msg = lib().mhe_last_error_string()
        raise MheError(f'{what} failed (status {status}): {msg.decode() if msg else "?"}')


def ptr(t: torch.Tensor | None) -> c_void_p:
    if t is None:
        return c_void_p(0)
    return c_void_p(t.data_ptr())


def stream_ptr(device=None) -> c_void_p:
    return c_void_p(torch.cuda.current_stream(device).cuda_stream)


def require_cuda_f32(*tensors: torch.Tensor) -> None:
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise MheError('mhentropy_b200 kernels need CUDA tensors (no CPU fallback on this path)')
        if t.dtype != torch.float32 or not t.is_contiguous():
            raise MheError(f'expected contiguous float32, got {t.dtype} contiguous={t.is_contiguous()}')


class Workspace:
    """Grow-only scratch buffer per device (the library never allocates)."""

    def __init__(self):
        self._buf: dict = {}

    def get(self, nbytes: int, device, tag: str = 'ws') -> torch.Tensor:
        key = (str(device), tag)
        buf = self._buf.get(key)
        if buf is None or buf.numel() < nbytes:
            buf = torch.empty(max(int(nbytes), 1024), dtype=torch.uint8, device=device)
            self._buf[key] = buf
        return buf


WORKSPACE = Workspace()
