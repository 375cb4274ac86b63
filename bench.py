#!/usr/bin/env python
"""Benchmark of the MHEntropy multi-hypothesis hot path on B200 (contract: see DESIGN.md §Measurement).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one fused training step from the image feature on, for B images x S hypotheses:
hoisted conditioning, flow sample + log q (entropy), z assembly, MANO, visible-2D reprojection + priors,
N-means, loss, and the full backward (flow weights, feature, det-head outputs).  N = 1 runs BASELINE.json
configs[1] (B=64, S=10); N > 1 keeps 64 images per GPU (weak scaling) and all-reduces the flat gradient and
the loss over NCCL inside the timed region.  Prints ONE JSON line on rank 0.

`--impl reference` times the reference's CPU PyTorch algorithm for the same step (the oracle port in
oracle/, two flow passes + autograd exactly as hand/network.py does) on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

OUT = sys.stdout
METRIC = 'hypotheses/sec (flow+MANO fwd+bwd)'
UNIT = 'hypotheses/s'

# algorithmic FLOP per hypothesis (SURVEY.md §8d / BASELINE.md §4)
FLOW_PASS_FLOP = 14_794_752
COND_FLOP_PER_IMAGE = 25_165_824
MANO_FWD_FLOP = 1_255_950


def train_step_flop_per_hyp(S: int) -> float:
    return 3 * (FLOW_PASS_FLOP + MANO_FWD_FLOP) + 3 * COND_FLOP_PER_IMAGE / S


def measured_peaks() -> dict:
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path) as fh:
            d = json.load(fh)
        return {'hbm_gbs': d['hbm_gbs'], 'bf16_tflops': d['bf16_tflops'], 'bf16_tflops_sustained': d['bf16_tflops_sustained'],
                'source': 'measured'}
    return {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0, 'bf16_tflops_sustained': 1400.0, 'source': 'fallback'}


class ClockSampler:
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index: int):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {
            getattr(nv, 'nvmlClocksEventReasonHwSlowdown', 0x8): 'hw_slowdown',
            getattr(nv, 'nvmlClocksEventReasonHwThermalSlowdown', 0x40): 'hw_thermal_slowdown',
            getattr(nv, 'nvmlClocksEventReasonSwThermalSlowdown', 0x20): 'sw_thermal_slowdown',
            getattr(nv, 'nvmlClocksEventReasonSwPowerCap', 0x4): 'sw_power_cap',
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.004)

    def __enter__(self):
        if self.nv is not None:
            self._thr = threading.Thread(target=self._loop, daemon=True)
            self._thr.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        if self._thr is not None:
            self._thr.join()

    def summary(self) -> dict:
        if not self.samples:
            return {'sm_mhz': None, 'sm_max_mhz': self.max_mhz, 'reasons': ['nvml unavailable']}
        return {'sm_mhz': statistics.median(self.samples), 'sm_max_mhz': self.max_mhz, 'reasons': sorted(self.reasons)}


# ---------------------------------------------------------------------------------------------------
def cpu_reference_run(B: int, S: int, steps: int, warmup: int, budget_s: float):
    """The reference's algorithm for the step on the host cores (oracle port, torch CPU, all threads)."""
    from mhentropy_b200.mano_assets import synthetic_mano
    from mhentropy_b200.synthetic import synthetic_batch
    from oracle import flow_oracle as fo, loss_oracle as lo, mano_oracle as mo

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = fo.init_state_dict(seed=0)
    sdg = {k: v.clone().requires_grad_(k != 'mask') for k, v in sd.items()}
    c = mo.mano_constants(synthetic_mano(0))
    batch = synthetic_batch(B, S, seed=0)

    def step():
        for v in sdg.values():
            v.grad = None
        feat = batch['feat'].clone().requires_grad_(True)
        zd = batch['z_det'].clone().requires_grad_(True)
        out = lo.reverse_kld(sdg, c, feat, zd, batch['z0'], batch['crop_uv'], batch['vis'], S)
        loss = lo.mhent_loss(out['log_p'])
        loss.backward()
        return float(loss)

    t_start = time.perf_counter()
    for _ in range(max(warmup, 1)):
        step()
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        step()
        times.append(time.perf_counter() - t0)
        if time.perf_counter() - t_start > budget_s and len(times) >= 3:
            break
    total = sum(times)
    return {'value': B * S * len(times) / total, 'steps': len(times), 'ms_per_step': 1e3 * total / len(times), 'cores': cores}


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    B, S = args.batch, args.hyp
    r = cpu_reference_run(B, S, args.steps, args.warmup, budget_s=150.0)
    sample = f'{r["steps"]} full steps of B={B} x S={S} (fwd+bwd), torch CPU fp32, {r["cores"]} threads'
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': r['value'], 'unit': UNIT, 'n_gpus': args.gpus, 'steps': r['steps'],
        'warmup': args.warmup, 'ms_per_step': r['ms_per_step'], 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': f'training step B={B} x S={S} hypotheses (BASELINE configs[1]): flow sample+log_prob, MANO, '
                               'visible-2D + entropy loss fwd+bwd', 'images_per_gpu': B, 'hypotheses': S},
        'cpu_baseline': {'value': r['value'], 'unit': UNIT, 'cores': r['cores'], 'kind': 'port', 'sample': sample},
        'e2e': {'value': r['value'], 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), file=OUT, flush=True)


# ---------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch.distributed as dist
    from mhentropy_b200 import MHEntHead, _lib
    from mhentropy_b200.engine import TrainStep
    from mhentropy_b200.mano_assets import synthetic_mano
    from mhentropy_b200.synthetic import synthetic_batch

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device (the product path has no CPU fallback); '
                         'use --impl reference for the CPU arm')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    B, S = args.batch, args.hyp
    R = B * S
    L = _lib.lib()

    torch.manual_seed(0)
    head = MHEntHead(mano_data=synthetic_mano(0)).to(dev)       # random-init flow weights (seed 0), synthetic MANO
    head.q_z_giv_i.precision = args.precision
    for p in head.parameters():
        p.requires_grad_(True)
    # every rank draws its own images; inputs start in pinned host memory for the e2e leg
    batch = synthetic_batch(B, S, seed=1000 + rank)
    host = {k: v.pin_memory() for k, v in batch.items()}
    devb = {k: v.to(dev) for k, v in batch.items()}
    h2d_bytes = sum(v.numel() * 4 for v in batch.values())

    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)   # > 126 MB L2

    # multi-GPU: ONE NCCL all-reduce of the flat gradient (+ the loss) after the step.  MHE_BENCH_ALLREDUCE_INSIDE=1 runs it inside the
    # engine's captured step instead, bucketed by backward chunk (mhe_flow_join_chunk) and overlapped with the remaining chunk: measured
    # slower on 2 GPUs (0.774 vs 0.743 ms/step) - inside the graph NCCL runs the 15-25 MB buckets with its LL protocol and its CTAs
    # compete with the cluster kernel for SMs - so it is not the default.
    ar_inside = world > 1 and bool(os.environ.get('MHE_BENCH_ALLREDUCE_INSIDE'))

    def allreduce(engine):
        if world > 1:
            engine.exchange_gradients()      # factored exchange (MHE_BENCH_DENSE_ALLREDUCE=1: one dense all-reduce of the 80 MB)

    # ---------------- value: inputs resident in HBM, fused engine (CUDA graph) ----------------
    eng = TrainStep(head, B, S, dev, want_verts=True, use_graph=not args.no_graph, allreduce=ar_inside,
                    factored_exchange=world > 1 and not os.environ.get('MHE_BENCH_DENSE_ALLREDUCE'))
    eng.load(**devb)
    for _ in range(max(args.warmup, 3)):
        flush.zero_()
        eng.run()
        allreduce(eng)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    launches0 = L.mhe_kernel_launch_count()
    with ClockSampler(local) as clocks:
        torch.cuda.synchronize()
        for a, b in ev:
            flush.zero_()                    # evict weights/activations from L2 between timed steps
            a.record()
            eng.run()
            allreduce(eng)
            b.record()
        torch.cuda.synchronize()
    step_ms = [a.elapsed_time(b) for a, b in ev]
    total_ms = torch.tensor([sum(step_ms)], device=dev)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_ms = float(total_ms)
    launches_per_step = eng.launches_per_step
    value = world * R * args.steps / (total_ms * 1e-3)

    if args.profile:
        if rank == 0:
            print(json.dumps({'profile_only': True, 'ms_per_step': total_ms / args.steps, 'gpu_launches_per_step': int(launches_per_step)}), file=OUT, flush=True)
        return

    # ---------------- e2e: public API with HOST buffers inside the timed region --------------------------------
    # (a) TrainStep.load(host tensors) + run() + loss.item(): pinned host inputs -> device, the step, loss -> host
    def e2e_step():
        eng.load(**host)
        eng.run()
        allreduce(eng)
        return eng.loss.item()               # device -> host read of the step's result

    # (b) the drop-in modules under autograd: MHEntHead.get_loss(...) + loss.backward()
    def e2e_autograd_step():
        feat = host['feat'].to(dev, non_blocking=True).requires_grad_(True)
        z_det = host['z_det'].to(dev, non_blocking=True).requires_grad_(True)
        z0 = host['z0'].to(dev, non_blocking=True)
        y = {'crop_uv': host['crop_uv'].to(dev, non_blocking=True), 'vis': host['vis'].to(dev, non_blocking=True)}
        head.zero_grad(set_to_none=True)
        out = head.get_loss(feat, y, z0=z0, z_det=z_det, N=S, want_verts=True)
        loss = (-out['log_p']).mean()
        loss.backward()
        if world > 1:
            dist.all_reduce(head.q_z_giv_i._last_flat_grad)
        return loss.item()

    def time_e2e(fn, steps):
        for _ in range(max(args.warmup, 3)):
            fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(steps):
            fn()
        b.record()
        torch.cuda.synchronize()
        ms = torch.tensor([a.elapsed_time(b)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return world * R * steps / (float(ms) * 1e-3)

    e2e_value = time_e2e(e2e_step, args.steps)
    e2e_autograd_value = time_e2e(e2e_autograd_step, max(args.steps // 2, 5))

    # ---------------- roofline of the dominant kernels, timed live with CUDA events on their launching stream -------
    # The two cluster-fused flow kernels (all 12 coupling layers forward; all 12 layers of data gradients backward) are ~60 % of
    # the step.  Algorithmic FLOPs per launch: one flow pass = 14,794,752 FLOP per row, both for the forward kernel
    # and for the backward kernel (which computes the data gradients; the weight gradients are separate batched GEMMs).
    peaks = measured_peaks()
    roof = None
    if rank == 0:
        probe_eng = TrainStep(head, B, S, dev, want_verts=True, use_graph=False)
        probe_eng.load(**devb)
        import ctypes

        PROBE_RUNS = 5

        def probe(tag):
            _lib.check(L.mhe_probe_configure(tag, 4096), 'probe')
            for _ in range(3):
                flush.zero_()
                probe_eng.run()
            torch.cuda.synchronize()
            L.mhe_probe_reset()
            for _ in range(PROBE_RUNS):
                flush.zero_()
                probe_eng.run()
            ms, n = ctypes.c_float(), ctypes.c_int()
            _lib.check(L.mhe_probe_read(ctypes.byref(ms), ctypes.byref(n)), 'probe read')
            return ms.value, n.value

        fused = args.precision == 'bf16x3'
        if fused:
            ms_f, n_f = probe(b'fused flow fwd')
            ms_b, n_b = probe(b'fused flow bwd')
            if n_f == 0 or n_b == 0:     # row count above the cluster-fused path (MHE_FUSED_MAX_ROWS): the per-GEMM tensor-core path ran
                fused = False
        if not fused:
            ms_f, n_f = probe(b'flow G1')
            ms_b, n_b = probe(b'dgrad G1')
        L.mhe_probe_configure(None, 0)
        flop_pass = float(FLOW_PASS_FLOP) * R if fused else 2.0 * R * 512 * 512 * 2
        # a pass may be cut into several launches of the same kernel (chunks of consecutive layers): FLOPs per launch scale with it
        lpp_f, lpp_b = max(n_f, 1) / PROBE_RUNS, max(n_b, 1) / PROBE_RUNS
        us_f, us_b = ms_f / max(n_f, 1) * 1e3, ms_b / max(n_b, 1) * 1e3
        # (fused: FLOW_PASS_FLOP covers a whole pass = lpp launches; per-GEMM path: flop_pass already is ONE G1 launch of both nets)
        fl_f, fl_b = (flop_pass / lpp_f, flop_pass / lpp_b) if fused else (flop_pass, flop_pass)
        ach_f, ach_b = fl_f / max(us_f, 1e-9) * 1e6 / 1e12, fl_b / max(us_b, 1e-9) * 1e6 / 1e12
        step_us = total_ms / args.steps * 1e3
        traffic = None
        tpath = os.path.join(ROOT, 'profiles', 'r1_fused_bwd_traffic.json')
        if fused and os.path.exists(tpath):
            with open(tpath) as fh:
                traffic = json.load(fh).get('dram_bytes_per_launch')
        roof = {'bound': 'tensor',
                'kernel': ('flow_bwd_fused_kernel (cluster-fused data-gradient pass over all 12 coupling layers: tcgen05 split-bf16x3, TMA '
                           'weight ring, DSMEM exchanges)' if fused else ('tc_gemm_kernel (tcgen05 split precision), dgrad G1 of one layer' if args.precision == 'bf16x3'
                                                             else 'sgemm_kernel (fp32 CUDA cores), dgrad G1')),
                'achieved': ach_b, 'peak': peaks['bf16_tflops_sustained'], 'unit': 'TFLOP/s', 'frac': ach_b / peaks['bf16_tflops_sustained'],
                'traffic': traffic, 'peak_source': f'{peaks["source"]} bf16 dense sustained',
                'launches_timed': n_b, 'avg_launch_us': us_b, 'launches_per_step': lpp_b, 'share_of_step': us_b * lpp_b / step_us,
                'algorithmic_flop_per_launch': fl_b,
                'second_kernel': {'kernel': 'flow_fwd_fused_kernel (cluster-fused sampling pass, same mapping)' if fused else 'flow G1',
                                  'achieved': ach_f, 'frac': ach_f / peaks['bf16_tflops_sustained'], 'avg_launch_us': us_f,
                                  'launches_per_step': lpp_f, 'share_of_step': us_f * lpp_f / step_us, 'launches_timed': n_f},
                'note': ('split precision issues 3 tensor-core products per fp32 product (hi*hi + hi*lo + lo*hi, as two instructions); judged '
                         'against the bf16 dense peak with the 1x algorithmic FLOP count, so 1/3 is the ceiling.  At 640 rows the kernel is bound by '
                         'the per-SM L2->SMEM ingest of the weights (320 KB per layer per CTA at ~50 B/clk) and by the exchange latency between '
                         'the 8 CTAs of a cluster, not by the tensor pipe (DESIGN.md §5)')}
        flop_step = train_step_flop_per_hyp(S) * R
        weight_bytes = 20_030_520 * 4
        roof_step = {'algorithmic_gflop': flop_step / 1e9, 't_tensor_us': flop_step / (peaks['bf16_tflops_sustained'] * 1e12) * 1e6,
                     't_hbm_us': (3 * weight_bytes + R * 9_336) / (peaks['hbm_gbs'] * 1e9) * 1e6,
                     'measured_us': step_us}
        roof_step['frac_of_governing'] = max(roof_step['t_tensor_us'], roof_step['t_hbm_us']) / roof_step['measured_us']

    # ---------------- CPU baseline (rank 0, N = 1 only): bounded sample on the host cores ----------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        r = cpu_reference_run(B, S, steps=10, warmup=2, budget_s=25.0)
        cpu = {'value': r['value'], 'unit': UNIT, 'cores': r['cores'], 'kind': 'port',
               'sample': f'{r["steps"]} full steps of B={B} x S={S} fwd+bwd (oracle port of the reference algorithm, torch CPU fp32)'}

    if rank == 0:
        line = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': max(args.warmup, 3),
            'ms_per_step': total_ms / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f32 (bf16x3 split on tcgen05, fp32 accumulate)' if args.precision == 'bf16x3' else 'f32', 'data': 'synthetic',
            'config': {'workload': f'training step B={B} x S={S} hypotheses per GPU (BASELINE configs[1]): flow sample+log_prob, '
                                   'MANO (778-vertex mesh fwd), visible-2D + entropy loss, fwd+bwd'
                                   + ((', gradient exchange: NCCL all-gather of the conditioning factors + local weight-gradient GEMM, '
                                       'NCCL all-reduce of the other 30 MB' if eng.factored_exchange else ', NCCL allreduce of the flat gradient')
                                      if world > 1 else ''),
                       'images_per_gpu': B, 'hypotheses': S, 'rows_per_gpu': R, 'l2': 'flushed (256 MiB write) before every timed step',
                       'launch': 'CUDA graph' if not args.no_graph else 'stream', 'parallelism': f'dp{world}'},
            'clocks': clocks.summary(),
            'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': h2d_bytes, 'd2h_bytes_per_step': 4,
                    'api': 'TrainStep.load(pinned host inputs) + TrainStep.run() + loss.item()',
                    'autograd_api_value': e2e_autograd_value,
                    'autograd_api': 'MHEntHead.get_loss + loss.backward() (drop-in modules), same host buffers'},
            'gpu_launches': int(launches_per_step) * args.steps,
            'gpu_launches_per_step': int(launches_per_step),
            'roofline': roof, 'step_roofline': roof_step, 'cpu_baseline': cpu,
        }
        print(json.dumps(line), file=OUT, flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def _claim_stdout():
    """Keep stdout for the ONE JSON line: libraries (NCCL's version banner, ...) that print to fd 1 are sent to stderr."""
    sys.stdout.flush()
    real = os.fdopen(os.dup(1), 'w')
    os.dup2(2, 1)
    return real


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=50)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--batch', type=int, default=64, help='images per GPU')
    ap.add_argument('--hyp', type=int, default=10, help='hypotheses per image')
    ap.add_argument('--no-graph', action='store_true')
    ap.add_argument('--precision', default='bf16x3', choices=['bf16x3', 'fp32'],
                    help="arithmetic of the flow contractions: tcgen05 split-bf16 (default) or exact fp32 CUDA cores")
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--profile', action='store_true', help='value leg only (for ncu launch lists)')
    args = ap.parse_args()
    global OUT
    OUT = _claim_stdout()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
