"""tcgen05/TMA GEMM building block against torch fp64 matmul: every operand-major combination, plain bf16
and bf16x3 split precision, both N tiles, ragged sizes and split-K."""
import ctypes

import pytest
import torch

from mhentropy_b200 import _lib

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def planes_of(x, rows_p, cols_p, planes, f16):
    """x fp32 [b][rows][cols] -> 16-bit planes (half or bfloat16) via the library's own converter."""
    b, rows, cols = x.shape
    out = torch.empty(b, planes, rows_p, cols_p, dtype=torch.bfloat16, device=DEV)
    _lib.check(_lib.lib().mhe_split_planes(_lib.ptr(x.contiguous()), rows, cols, _lib.ptr(out), rows_p, cols_p, planes, b, int(f16),
                                           _lib.stream_ptr()), 'split')
    return out


def run(M, N, K, batches, planes, a_mn, b_mn, bn, ksplit=1, seed=0, f16=False):
    """All extents are multiples of 8 (TMA strides are 16-byte multiples), so no padding is involved."""
    assert M % 8 == 0 and N % 8 == 0 and K % 8 == 0
    g = torch.Generator(device='cpu').manual_seed(seed)
    A = torch.randn(batches, M, K, generator=g).to(DEV)
    B = torch.randn(batches, N, K, generator=g).to(DEV) * 0.05
    Ap = planes_of(A.transpose(1, 2) if a_mn else A, K if a_mn else M, M if a_mn else K, planes, f16)
    Bp = planes_of(B.transpose(1, 2) if b_mn else B, K if b_mn else N, N if b_mn else K, planes, f16)
    C = torch.full((batches, M, N), float('nan'), device=DEV)
    st = _lib.lib().mhe_tc_gemm_raw(_lib.ptr(Ap), _lib.ptr(Bp), _lib.ptr(C), M, N, K, batches, planes, int(a_mn), int(b_mn), bn, ksplit,
                                    int(f16), _lib.stream_ptr())
    _lib.check(st, 'tc_gemm_raw')
    torch.cuda.synchronize()
    ref = A.double() @ B.double().transpose(1, 2)
    return float((C.double() - ref).abs().max() / ref.abs().max())


@pytest.mark.parametrize('a_mn,b_mn', [(False, False), (False, True), (True, False), (True, True)])
@pytest.mark.parametrize('bn', [64, 128])
def test_bf16x3_matches_fp64(a_mn, b_mn, bn):
    err = run(256, 256, 256, 2, 2, a_mn, b_mn, bn)
    assert err < 2e-5, err


@pytest.mark.parametrize('a_mn,b_mn', [(False, False), (True, True), (False, True), (True, False)])
def test_plain_bf16(a_mn, b_mn):
    err = run(128, 128, 128, 1, 1, a_mn, b_mn, 128)
    assert err < 2e-2, err


@pytest.mark.parametrize('M,N,K', [(640, 512, 512), (640, 64, 512), (200, 72, 136), (64, 512, 64), (136, 200, 1000)])
def test_ragged_and_flow_shapes(M, N, K):
    assert run(M, N, K, 2, 2, False, False, 64) < 2e-5
    assert run(M, N, K, 1, 2, True, True, 128) < 2e-5


@pytest.mark.parametrize('a_mn,b_mn', [(False, False), (False, True), (True, True)])
def test_f16x3_is_fp32_accurate(a_mn, b_mn):
    """Half planes keep ~22 significant bits: the three-pass product is as accurate as an fp32 GEMM."""
    err = run(256, 256, 512, 2, 2, a_mn, b_mn, 64, f16=True)
    assert err < 5e-6, err      # bf16x3 on the same data: ~1e-5


def test_split_k_atomic():
    assert run(128, 256, 2048, 1, 2, False, True, 128, ksplit=8) < 2e-5
    assert run(128, 128, 640, 2, 2, True, True, 64, ksplit=3) < 2e-5
