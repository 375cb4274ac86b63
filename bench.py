#!/usr/bin/env python
"""Benchmark of the MHEntropy multi-hypothesis hot path on B200 (contract: see DESIGN.md §Measurement).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one fused training step from the image feature on, for B images x S hypotheses:
hoisted conditioning, flow sample + log q (entropy), z assembly, MANO, visible-2D reprojection + priors,
N-means, loss, and the full backward (flow weights, feature, det-head outputs).  N = 1 runs BASELINE.json
configs[1] (B=64, S=10); N > 1 keeps 64 images per GPU (weak scaling) and exchanges the flat gradient and
the loss over NCCL inside the timed region.  Prints ONE JSON line on rank 0.

The same line carries, under "configs", the other BASELINE.json configurations as driver-observed
numbers: config 1 (B=8 x S=10, the reference's CPU case, oracle port on the host cores), config 3
(inference sampling 256 x 100 with meshes), config 4 (NLL scoring of 16,384 poses), the per-GPU shard of
config 5 (512 x 64) and config 5 itself (4096 x 64 split over the N ranks: strong scaling).

`--impl reference` times the reference's CPU PyTorch algorithm for the same step (the oracle port in
oracle/, two flow passes + autograd exactly as hand/network.py does) on the host cores.
"""
from __future__ import annotations

import argparse
import hashlib
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


def workload(B: int, S: int) -> str:
    """The ONE description of the measured workload, shared verbatim by both arms."""
    return (f'training step B={B} x S={S} hypotheses per GPU (BASELINE configs[1]): flow sample+log_prob, MANO, '
            'visible-2D + entropy loss, fwd+bwd')


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
def cpu_reference_run(B: int, S: int, steps: int, warmup: int, budget_s: float, backward: bool = True):
    """The reference's algorithm for the step on the host cores (oracle port, torch CPU, all threads)."""
    from mhentropy_b200.mano_assets import synthetic_mano
    from mhentropy_b200.synthetic import synthetic_batch
    from oracle import flow_oracle as fo, loss_oracle as lo, mano_oracle as mo

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = fo.init_state_dict(seed=0)
    sdg = {k: v.clone().requires_grad_(k != 'mask' and backward) for k, v in sd.items()}
    c = mo.mano_constants(synthetic_mano(0))
    batch = synthetic_batch(B, S, seed=0)

    def step():
        for v in sdg.values():
            v.grad = None
        with torch.set_grad_enabled(backward):
            feat = batch['feat'].clone().requires_grad_(backward)
            zd = batch['z_det'].clone().requires_grad_(backward)
            out = lo.reverse_kld(sdg, c, feat, zd, batch['z0'], batch['crop_uv'], batch['vis'], S)
            loss = lo.mhent_loss(out['log_p'])
            if backward:
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
    return {'value': B * S * len(times) / total, 'steps': len(times), 'ms_per_step': 1e3 * total / len(times),
            'ms_median': 1e3 * statistics.median(times), 'cores': cores}


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
        'config': {'workload': workload(B, S), 'images_per_gpu': B, 'hypotheses': S,
                   'detail': 'oracle port of the reference step (two flow passes + autograd, hand/network.py:760-831) on the host cores'},
        'cpu_baseline': {'value': r['value'], 'unit': UNIT, 'cores': r['cores'], 'kind': 'port', 'sample': sample},
        'e2e': {'value': r['value'], 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), file=OUT, flush=True)


# ---------------------------------------------------------------------------------------------------
def _sha(path: str) -> str:
    with open(path, 'rb') as fh:
        return hashlib.sha256(fh.read()).hexdigest()[:16]


def committed_traffic():
    """DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture - only while the capture was taken
    from the kernel source that is in the tree now (a stale figure is worse than none)."""
    path = os.path.join(ROOT, 'profiles', 'r2_fused_bwd_traffic.json')
    if not os.path.exists(path):
        return None
    with open(path) as fh:
        d = json.load(fh)
    src = os.path.join(ROOT, 'mhentropy_b200', 'csrc', 'flow_fused.cu')
    if d.get('flow_fused_cu_sha256_16') != _sha(src):
        return None
    return d.get('dram_bytes_per_launch')


def run_ours(args):
    import torch.distributed as dist
    from mhentropy_b200 import FlatAdam, MHEntHead, _lib
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
    peaks = measured_peaks()

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

    # multi-GPU: TrainStep.exchange_gradients() after the step - factored (all-gather of the conditioning factors + local weight-gradient
    # GEMM, all-reduce of the other 30 MB) while the factors are well below the dense gradient (2, 4 ranks), one dense all-reduce above
    # (8 ranks); MHE_BENCH_EXCHANGE=dense|factored forces one.  MHE_BENCH_ALLREDUCE_INSIDE=1: bucketed all-reduce inside the captured step.
    exchange = os.environ.get('MHE_BENCH_EXCHANGE', 'dense' if os.environ.get('MHE_BENCH_DENSE_ALLREDUCE') else 'auto')
    ar_inside = world > 1 and (bool(os.environ.get('MHE_BENCH_ALLREDUCE_INSIDE')) or exchange == 'inside')
    if exchange == 'inside':
        exchange = 'dense'

    def allreduce(engine):
        if world > 1:
            engine.exchange_gradients()

    def timed_steps(engine, steps, warmup):
        """Device-timed steps with the L2 flushed before each: (per-step ms list, max-over-ranks total ms)."""
        for _ in range(max(warmup, 3)):
            flush.zero_()
            engine.run()
            allreduce(engine)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        torch.cuda.synchronize()
        for a, b in ev:
            flush.zero_()                    # evict weights/activations from L2 between timed steps
            a.record()
            engine.run()
            allreduce(engine)
            b.record()
        torch.cuda.synchronize()
        ms = [a.elapsed_time(b) for a, b in ev]
        tot = torch.tensor([sum(ms)], device=dev)
        if world > 1:
            dist.all_reduce(tot, op=dist.ReduceOp.MAX)
        return ms, float(tot)

    # ---------------- value: inputs resident in HBM, fused engine (CUDA graph) ----------------
    # The split weight planes come from the optimizer (FlatAdam.step refreshes them in the same pass as the update), so
    # no conversion kernel runs inside the step; `step_variants` below also times the step that re-packs them itself and step + Adam.
    eng = TrainStep(head, B, S, dev, want_verts=True, use_graph=not args.no_graph, allreduce=ar_inside, planes=args.planes, exchange=exchange)
    eng.load(**devb)
    launches0 = L.mhe_kernel_launch_count()
    with ClockSampler(local) as clocks:
        step_ms, total_ms = timed_steps(eng, args.steps, args.warmup)
    launches_per_step = eng.launches_per_step
    value = world * R * args.steps / (total_ms * 1e-3)
    factored = eng.factored_exchange
    peer_on = eng.px is not None
    del launches0

    if args.profile:
        if rank == 0:
            print(json.dumps({'profile_only': True, 'ms_per_step': total_ms / args.steps, 'gpu_launches_per_step': int(launches_per_step)}), file=OUT, flush=True)
        return

    # ---------------- e2e: public API with HOST buffers inside the timed region --------------------------------
    # (a) TrainStep.load(host tensors) + run() + loss.item(): pinned host inputs -> device, the step, loss -> host
    stage = eng.staging()                    # the engine's pinned staging block: a loader collates into these views
    h2d_bytes = eng.inputs.numel() * 4       # what load_staged() copies (the five input tensors; segments padded to 128 bytes)
    for k, v in host.items():
        stage[k].copy_(v)

    def e2e_step():
        eng.load_staged()                    # ONE host-to-device copy of the step's inputs (h2d_bytes_per_step counts the tensors)
        eng.run()
        allreduce(eng)
        return eng.loss.item()               # device -> host read of the step's result

    # (b) the drop-in modules under autograd: MHEntHead.get_loss(...) + loss.backward()
    flow = head.q_z_giv_i
    plist = list(head.parameters())      # what an optimizer holds: zeroing walks this flat list (optimizer.zero_grad(), as the reference's loop
                                         # does, CrossModalHand.py:454), not the module tree (Module.zero_grad costs ~0.3 ms of host time here)

    def e2e_autograd_step():
        feat = host['feat'].to(dev, non_blocking=True).requires_grad_(True)
        z_det = host['z_det'].to(dev, non_blocking=True).requires_grad_(True)
        z0 = host['z0'].to(dev, non_blocking=True)
        y = {'crop_uv': host['crop_uv'].to(dev, non_blocking=True), 'vis': host['vis'].to(dev, non_blocking=True)}
        for p in plist:
            p.grad = None
        out = head.get_loss(feat, y, z0=z0, z_det=z_det, N=S, want_verts=True)
        loss = (-out['log_p']).mean()
        loss.backward()
        if world > 1:
            dist.all_reduce(flow.grad_buffer())          # every parameter's .grad is a view of this one buffer
        return loss.item()

    def time_e2e(fn, steps, flushed):
        for _ in range(max(args.warmup, 3)):
            fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        if flushed:      # same cache state as `value`: L2 flushed before every step, each step timed with its own event pair
            tot = 0.0
            for _ in range(steps):
                flush.zero_()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                fn()
                b.record()
                torch.cuda.synchronize()
                tot += a.elapsed_time(b)
        else:
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(steps):
                fn()
            b.record()
            torch.cuda.synchronize()
            tot = a.elapsed_time(b)
        ms = torch.tensor([tot], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return world * R * steps / (float(ms) * 1e-3)

    e2e_value = time_e2e(e2e_step, args.steps, True)
    e2e_noflush = time_e2e(e2e_step, args.steps, False)
    e2e_autograd_value = time_e2e(e2e_autograd_step, max(args.steps // 2, 5), True)

    # ---------------- step variants (N = 1): planes re-packed inside the step; step + optimizer ----------------
    variants = None
    if world == 1 and rank == 0:
        variants = {}
        other = 'step' if args.planes == 'optimizer' else 'optimizer'
        eng2 = TrainStep(head, B, S, dev, want_verts=True, use_graph=not args.no_graph, planes=other)
        eng2.load(**devb)
        _, tot2 = timed_steps(eng2, max(args.steps // 2, 5), 3)
        variants[f'planes_{other}_ms_per_step'] = tot2 / max(args.steps // 2, 5)
        del eng2
        if args.precision == 'bf16x3':
            # the reference's update (clip_grad_norm_ + Adam, CrossModalHand.py:462-470) on the flat buffer with the plane refresh:
            # what a training loop runs between two steps when planes == 'optimizer'.  lr = 0: the weights stay put for the later legs.
            opt = FlatAdam(flow, lr=0.0)
            n = max(args.steps // 2, 5)
            for _ in range(3):
                eng.run()
                opt.step(eng.dflat)
            torch.cuda.synchronize()
            tot = 0.0
            for _ in range(n):
                flush.zero_()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                eng.run()
                opt.grad_sqnorm(eng.dflat)
                opt.step(eng.dflat)
                b.record()
                torch.cuda.synchronize()
                tot += a.elapsed_time(b)
            variants['step_plus_clipnorm_adam_plane_refresh_ms'] = tot / n

    # ---------------- roofline of the dominant kernels, timed live with CUDA events on their launching stream -------
    # The two cluster-fused flow kernels (all 12 coupling layers forward; all 12 layers of data gradients backward) are ~70 % of
    # the step.  Algorithmic FLOPs per launch: one flow pass = 14,794,752 FLOP per row, both for the forward kernel
    # and for the backward kernel (which computes the data gradients; the weight gradients are separate batched GEMMs).
    roof = roof_step = None
    if rank == 0:
        probe_eng = TrainStep(head, B, S, dev, want_verts=True, use_graph=False, planes=args.planes)
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
        del probe_eng
        flop_pass = float(FLOW_PASS_FLOP) * R if fused else 2.0 * R * 512 * 512 * 2
        # a pass may be cut into several launches of the same kernel (chunks of consecutive layers): FLOPs per launch scale with it
        lpp_f, lpp_b = max(n_f, 1) / PROBE_RUNS, max(n_b, 1) / PROBE_RUNS
        us_f, us_b = ms_f / max(n_f, 1) * 1e3, ms_b / max(n_b, 1) * 1e3
        # (fused: FLOW_PASS_FLOP covers a whole pass = lpp launches; per-GEMM path: flop_pass already is ONE G1 launch of both nets)
        fl_f, fl_b = (flop_pass / lpp_f, flop_pass / lpp_b) if fused else (flop_pass, flop_pass)
        ach_f, ach_b = fl_f / max(us_f, 1e-9) * 1e6 / 1e12, fl_b / max(us_b, 1e-9) * 1e6 / 1e12
        step_us = total_ms / args.steps * 1e3
        roof = {'bound': 'tensor',
                'kernel': ('flow_bwd_fused_kernel (cluster-fused data-gradient pass over all 12 coupling layers: tcgen05 split-bf16x3, TMA '
                           'weight ring, DSMEM exchanges)' if fused else ('tc_gemm_kernel (tcgen05 split precision), dgrad G1 of one layer' if args.precision == 'bf16x3'
                                                             else 'sgemm_kernel (fp32 CUDA cores), dgrad G1')),
                'achieved': ach_b, 'peak': peaks['bf16_tflops_sustained'], 'unit': 'TFLOP/s', 'frac': ach_b / peaks['bf16_tflops_sustained'],
                'traffic': committed_traffic() if fused else None, 'peak_source': f'{peaks["source"]} bf16 dense sustained',
                'launches_timed': n_b, 'avg_launch_us': us_b, 'launches_per_step': lpp_b, 'share_of_step': us_b * lpp_b / step_us,
                'algorithmic_flop_per_launch': fl_b,
                'second_kernel': {'kernel': 'flow_fwd_fused_kernel (cluster-fused sampling pass, same mapping)' if fused else 'flow G1',
                                  'achieved': ach_f, 'frac': ach_f / peaks['bf16_tflops_sustained'], 'avg_launch_us': us_f,
                                  'launches_per_step': lpp_f, 'share_of_step': us_f * lpp_f / step_us, 'launches_timed': n_f},
                'note': ('split precision issues 3 tensor-core products per fp32 product (hi*hi + hi*lo + lo*hi, as two instructions); judged '
                         'against the bf16 dense peak with the 1x algorithmic FLOP count, so 1/3 is the ceiling (profiles/r2_precision_schemes.json: '
                         'every 2-product / tf32 variant misses the 1e-3 gradient bar).  At 640 rows the kernel is bound by the per-SM L2->SMEM '
                         'ingest of the weights (~416 KB per layer per CTA) and by the exchange latency between the 8 CTAs of a cluster, not by '
                         'the tensor pipe (DESIGN.md §5)')}
        flop_step = train_step_flop_per_hyp(S) * R
        weight_bytes = 20_030_520 * 4
        roof_step = {'algorithmic_gflop': flop_step / 1e9, 't_tensor_us': flop_step / (peaks['bf16_tflops_sustained'] * 1e12) * 1e6,
                     't_hbm_us': (3 * weight_bytes + R * 9_336) / (peaks['hbm_gbs'] * 1e9) * 1e6,
                     'measured_us': step_us}
        roof_step['frac_of_governing'] = max(roof_step['t_tensor_us'], roof_step['t_hbm_us']) / roof_step['measured_us']

    # ---------------- the other BASELINE.json configurations, driver-observed in the same line ----------------
    configs = {}
    if not args.no_configs:
        del eng
        torch.cuda.empty_cache()
        configs = other_configs(args, head, dev, flush, rank, world, peaks)

    # ---------------- CPU baseline (rank 0, N = 1 only): bounded sample on the host cores ----------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        r = cpu_reference_run(B, S, steps=10, warmup=2, budget_s=20.0)
        cpu = {'value': r['value'], 'unit': UNIT, 'cores': r['cores'], 'kind': 'port',
               'sample': f'{r["steps"]} full steps of B={B} x S={S} fwd+bwd (oracle port of the reference algorithm, torch CPU fp32)'}
        # BASELINE.json configs[0] / BASELINE.md section 3: the reference's own CPU-runnable case, forward-only and forward+backward
        r1 = cpu_reference_run(8, 10, steps=10, warmup=3, budget_s=8.0)
        r1f = cpu_reference_run(8, 10, steps=10, warmup=3, budget_s=4.0, backward=False)
        configs['config1_cpu'] = {'workload': 'BASELINE configs[0]: hand flow+MANO forward + entropy loss, B=8 x S=10, fp32 on the host cores (oracle port)',
                                  'cores': r1['cores'], 'fwd_bwd_ms_median': r1['ms_median'], 'fwd_bwd_hyp_per_s': 80 / (r1['ms_median'] * 1e-3),
                                  'fwd_ms_median': r1f['ms_median'], 'fwd_hyp_per_s': 80 / (r1f['ms_median'] * 1e-3), 'steps': r1['steps'],
                                  'algorithmic_gflop_fwd_bwd': 80 * train_step_flop_per_hyp(10) / 1e9}

    if rank == 0:
        exch = None
        if world > 1:
            exch = ('peer-memory exchange inside the captured step: copy-engine pushes over NVLink per backward chunk (reduce-scatter, '
                    'mhe_sum_shards, all-gather), no NCCL kernel' if peer_on else
                    'bucketed NCCL all-reduce inside the captured step' if ar_inside else
                    ('NCCL all-gather of the conditioning factors + local weight-gradient GEMM, NCCL all-reduce of the other 30 MB'
                     if factored else 'one NCCL all-reduce of the flat 80 MB gradient + loss'))
        line = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': max(args.warmup, 3),
            'ms_per_step': total_ms / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f32 (bf16x3 split on tcgen05, fp32 accumulate)' if args.precision == 'bf16x3' else 'f32', 'data': 'synthetic',
            'config': {'workload': workload(B, S), 'images_per_gpu': B, 'hypotheses': S, 'rows_per_gpu': R,
                       'detail': '778-vertex mesh forward included; split weight planes '
                                 + ('supplied by the optimizer (FlatAdam refreshes them with the update)' if args.planes == 'optimizer' else 're-packed inside the step'),
                       'gradient_exchange': exch, 'gradients': 'global-batch mean (each rank seeds 1/(B*world), the exchange sums)',
                       'l2': 'flushed (256 MiB write) before every timed step, in the value AND the e2e leg',
                       'launch': 'CUDA graph' if not args.no_graph else 'stream', 'parallelism': f'dp{world}'},
            'clocks': clocks.summary(),
            'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': h2d_bytes, 'd2h_bytes_per_step': 4,
                    'api': 'TrainStep.staging() pinned block -> load_staged() (one H2D copy) + TrainStep.run() + loss.item()',
                    'value_without_l2_flush': e2e_noflush,
                    'autograd_api_value': e2e_autograd_value,
                    'autograd_api': 'MHEntHead.get_loss + loss.backward() (drop-in modules), same host buffers, L2 flushed'},
            'gpu_launches': int(launches_per_step) * args.steps,
            'gpu_launches_per_step': int(launches_per_step),
            'roofline': roof, 'step_roofline': roof_step, 'step_variants': variants, 'cpu_baseline': cpu, 'configs': configs,
        }
        print(json.dumps(line), file=OUT, flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def other_configs(args, head, dev, flush, rank, world, peaks) -> dict:
    """BASELINE.json configs[2], [3], [4] as driver-observed numbers (each entry: ms, units/s, algorithmic GFLOP, fraction of the bf16
    sustained peak with the 1x FLOP count).  Configs 3 and 4 have no collective (replicas): every rank runs them, rank 0 reports its own
    time and value = world x that.  Config 5 splits its 4096 images over the ranks (strong scaling) and exchanges gradients."""
    import torch.distributed as dist
    from mhentropy_b200.engine import TrainStep
    from mhentropy_b200.synthetic import synthetic_batch

    peak = peaks['bf16_tflops_sustained'] * 1e12
    out = {}

    def timeit(fn, n, warm=2):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        tot = 0.0
        for _ in range(n):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize()
            tot += a.elapsed_time(b)
        t = torch.tensor([tot / n], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    def guarded(name, fn):
        try:
            out[name] = fn()
        except Exception as e:      # a secondary configuration must not take the headline line down with it
            out[name] = {'error': f'{type(e).__name__}: {e}'[:300]}
            torch.cuda.empty_cache()

    g = torch.Generator().manual_seed(3)

    def config3():
        B, S = 256, 100
        feat = torch.randn(B, 512, generator=g).to(dev)
        z0 = (torch.randn(B * S, 45, generator=g) * 0.8).to(dev)
        z_det = torch.cat([0.5 * torch.randn(B, 3, generator=g), 0.02 * torch.randn(B, 10, generator=g),
                           torch.randn(B, 1, generator=g) * 0.1 - 1.2, 0.1 * torch.randn(B, 2, generator=g)], 1).to(dev)
        res = {}

        def sample():
            res['o'] = head.sample(feat, N=S, temp=0.8, mods={'uv', 'xyz', 'verts'}, z0=z0, z_det=z_det)

        ms = timeit(sample, 5)
        assert all(torch.isfinite(v).all() for v in res['o'].values() if torch.is_tensor(v) and v.is_floating_point())
        flop = B * S * (FLOW_PASS_FLOP + MANO_FWD_FLOP) + B * COND_FLOP_PER_IMAGE
        return {'workload': 'BASELINE configs[2]: inference sampling B=256 x S=100 (flow sample + MANO mesh + reprojection), MHEntHead.sample',
                'rows_per_gpu': B * S, 'ms': ms, 'value': world * B * S / (ms * 1e-3), 'unit': UNIT, 'scaling': 'replicas (no collective)',
                'algorithmic_gflop': flop / 1e9, 'frac_of_bf16_sustained': flop / (ms * 1e-3) / peak}

    def config4():
        Rn = 16384
        x = (0.5 * torch.randn(Rn, 45, generator=g)).to(dev)
        featr = torch.randn(Rn, 512, generator=g).to(dev)
        res = {}

        def nll():
            with torch.no_grad():
                res['lp'] = head.q_z_giv_i.log_prob(x, logvar=featr)

        ms = timeit(nll, 5)
        assert torch.isfinite(res['lp']).all()
        flop = Rn * (FLOW_PASS_FLOP + COND_FLOP_PER_IMAGE)
        return {'workload': 'BASELINE configs[3]: NLL / log_prob scoring of 16,384 poses, one feature row per pose, flow inverse direction only',
                'rows_per_gpu': Rn, 'ms': ms, 'value': world * Rn / (ms * 1e-3), 'unit': 'poses/s', 'scaling': 'replicas (no collective)',
                'algorithmic_gflop': flop / 1e9, 'frac_of_bf16_sustained': flop / (ms * 1e-3) / peak}

    exchange_env = os.environ.get('MHE_BENCH_EXCHANGE', 'auto')

    def config5(images_total, tag):
        S = 64
        Bl = images_total // world
        batch = {k: v.to(dev) for k, v in synthetic_batch(Bl, S, seed=2000 + rank).items()}
        eng = TrainStep(head, Bl, S, dev, want_verts=True, use_graph=True, planes=args.planes,
                        exchange={'peer': 'auto', 'inside': 'dense'}.get(exchange_env, exchange_env))
        eng.load(**batch)

        def step():
            eng.run()
            if world > 1:
                eng.exchange_gradients()

        ms = timeit(step, 3, warm=2)
        assert torch.isfinite(eng.loss).all()
        flop = Bl * S * train_step_flop_per_hyp(S)
        r = {'workload': tag, 'images_total': Bl * world, 'images_per_gpu': Bl, 'hypotheses': S, 'rows_per_gpu': Bl * S, 'ms': ms,
             'value': world * Bl * S / (ms * 1e-3), 'unit': UNIT, 'launches_per_step': int(eng.launches_per_step),
             'algorithmic_gflop_per_gpu': flop / 1e9, 'frac_of_bf16_sustained': flop / (ms * 1e-3) / peak}
        del eng
        torch.cuda.empty_cache()
        return r

    def config1_gpu():
        Bs, Ss = 8, 10
        batch = {k: v.to(dev) for k, v in synthetic_batch(Bs, Ss, seed=3000 + rank).items()}
        eng = TrainStep(head, Bs, Ss, dev, want_verts=True, use_graph=True, planes=args.planes, exchange='dense')
        eng.load(**batch)

        def step():
            eng.run()
            if world > 1:
                eng.exchange_gradients()

        ms = timeit(step, 10, warm=3)
        assert torch.isfinite(eng.loss).all()
        r = {'workload': "BASELINE configs[0] shape on the GPU (the reference's CPU-runnable case; its CPU number is config1_cpu): training step "
                         'B=8 x S=10 per GPU', 'rows_per_gpu': Bs * Ss, 'ms': ms, 'value': world * Bs * Ss / (ms * 1e-3), 'unit': UNIT,
             'launches_per_step': int(eng.launches_per_step)}
        del eng
        torch.cuda.empty_cache()
        return r

    guarded('config1_gpu', config1_gpu)
    guarded('config3', config3)
    guarded('config4', config4)
    if world == 1:
        guarded('config5_shard', lambda: config5(512, 'per-GPU shard of BASELINE configs[4] at 8 GPUs: training step B=512 x S=64 on one GPU'))
    guarded('config5', lambda: {**config5(4096, 'BASELINE configs[4]: sharded training step B=4096 x S=64, images split over the ranks'),
                                'scaling': 'strong'})
    return out


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
    ap.add_argument('--planes', default='optimizer', choices=['optimizer', 'step'],
                    help="split weight planes supplied by the optimizer (FlatAdam's in-pass refresh) or re-packed inside every step")
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-configs', action='store_true', help='skip the secondary BASELINE configurations (configs 1, 3, 4, 5)')
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
