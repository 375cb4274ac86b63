"""Micro-benchmark of the tcgen05 GEMM building block on the coupling-layer shapes (L2-hot, back-to-back launches)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mhentropy_b200 import _lib

DEV = 'cuda'
L = _lib.lib()


def planes(b, rows, cols):
    return torch.randn(b, 2, rows, cols, device=DEV).bfloat16().contiguous()


def bench(name, M, N, K, b, a_mn, b_mn, bn, ksplit=1, reps=40):
    A = planes(b, K if a_mn else M, M if a_mn else K)
    B = planes(b, K if b_mn else N, N if b_mn else K)
    C = torch.zeros(b, M, N, device=DEV)
    s = _lib.stream_ptr()
    call = lambda: _lib.check(L.mhe_tc_gemm_raw(_lib.ptr(A), _lib.ptr(B), _lib.ptr(C), M, N, K, b, 2, int(a_mn), int(b_mn), bn, ksplit, 0, _lib.stream_ptr()), 'gemm')
    for _ in range(3):
        call()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            call()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / reps
    flops = 2.0 * M * N * K * b
    print(f'{name:28s} M={M:6d} N={N:4d} K={K:5d} b={b} bn={bn:3d} ks={ksplit}  {us:8.2f} us/launch  {flops / us / 1e6:8.1f} TFLOP/s (1x count)')


if __name__ == '__main__':
    if os.environ.get('G1_ONLY'):      # the long-batch G1 shape alone (MHE_RAW_NULL_EPI=1: main loop without the epilogue's stores)
        for R in (8192, 32768):
            bench('fwd G1 (K,K)', R, 512, 512, 2, False, False, 128)
            bench('dgrad G1 (K,MN)', R, 512, 512, 2, False, True, 128)
        sys.exit(0)
    for R in (640, 8192, 65536):
        print('--- rows', R)
        for bn in (64, 128):
            bench('fwd G1 (K,K)', R, 512, 512, 2, False, False, bn)
            bench('dgrad G1 (K,MN)', R, 512, 512, 2, False, True, bn)
            bench('wgrad W1 (MN,MN)', 512, 512, R, 2, True, True, bn, ksplit=1 if R < 8192 else 8)
        bench('fwd G0 K=64', R, 512, 64, 2, False, False, 64)
        bench('fwd G2 N=64', R, 64, 512, 2, False, False, 64)
        bench('wgrad W0 N=64 (MN,MN)', 512, 64, R, 2, True, True, 64, ksplit=1 if R < 8192 else 8)
