"""Fused multi-hypothesis training / sampling steps driven straight through the C ABI.

``TrainStep`` runs what ``MHEnt.get_loss`` + ``MHEntLoss`` + ``backward()`` run in the reference
(``hand/network.py:760-831``, ``hand/criteria.py:55,173``; call stack SURVEY.md §3.1) from the image feature
on: hoisted conditioning, ONE flow pass giving samples and log q (entropy), z assembly, MANO, visible-2D
reprojection + priors, N-means, and the full backward into a flat gradient buffer — a fixed sequence of
kernel launches on preallocated buffers, so it can be captured in a CUDA graph.  It computes exactly what
``MHEntHead.get_loss(...)`` + autograd computes (tests/test_gpu_engine.py).
"""
from __future__ import annotations

import contextlib
import os

import torch

from . import _lib
from ._lib import check, lib, ptr
from .losses import MHEntHead


def _env_on(name: str) -> bool:
    """An environment switch: set and neither empty nor '0'."""
    return os.environ.get(name, '') not in ('', '0')


def factored_exchange_pays(shape, B: int, world: int) -> bool:
    """Factored vs dense exchange of the conditioning weight gradient: the factors all-gathered per rank are
    world x B x (L*4*H + C) floats, the dense gradient L*4*H*C.  The gather also feeds a contraction over world x B images on every
    rank, so it only pays while the factors are a small fraction of the dense size.  Measured at 64 images per rank (ms/step,
    profiles/r2_exchange_modes_*): 2 ranks 0.655 factored vs 0.73 dense; 4 ranks 0.767 vs 0.752; 8 ranks 0.849 vs 0.842."""
    factors = world * B * (shape.layers * 4 * shape.hidden + shape.cond)
    dense = shape.layers * 4 * shape.hidden * shape.cond
    return factors <= 0.3 * dense


@contextlib.contextmanager
def _nvtx(name: str):
    """NVTX range around a phase of the step (shows up in Nsight Systems / ncu --nvtx; free when no tool is attached)."""
    torch.cuda.nvtx.range_push(name)
    try:
        yield
    finally:
        torch.cuda.nvtx.range_pop()


class TrainStep:
    """planes: ``'step'`` re-packs the split weight planes from the fp32 parameters inside every step (weights may have been changed by
    anything); ``'optimizer'`` reads the flow's own plane set (``RealNVP.packed_weights``), which ``FlatAdam.step`` refreshes in the
    same pass as the update - no conversion kernel runs inside the step.
    exchange: ``'auto'`` picks the factored or the dense gradient exchange by size (:func:`factored_exchange_pays`).
    average_over_ranks: with several ranks the loss and every gradient are the GLOBAL-batch mean (each rank seeds its backward with
    1 / (B * world) and the exchange sums), so a sharded step equals the single-process step on the concatenated batch."""

    def __init__(self, head: MHEntHead, B: int, S: int, device, want_verts: bool = True, use_graph: bool = True,
                 prepare_ahead: bool = False, pipelined_cond_bwd: bool = False, allreduce_group=None, allreduce: bool = False,
                 factored_exchange: bool | None = None, exchange_in_graph: bool = False, planes: str = 'step',
                 exchange: str = 'auto', average_over_ranks: bool = True):
        self.head, self.B, self.S, self.R = head, B, S, B * S
        self.dev = torch.device(device)
        flow = head.q_z_giv_i
        dist_on = torch.distributed.is_available() and torch.distributed.is_initialized()
        self.world = torch.distributed.get_world_size(allreduce_group) if dist_on else 1
        self.dloss = 1.0 / self.world if average_over_ranks else 1.0
        if planes not in ('step', 'optimizer'):
            raise ValueError("planes must be 'step' or 'optimizer'")
        self.planes = planes
        self.shape = flow._shape
        self.flat = flow.flat_parameters(self.dev)
        self.mask = flow.mask
        self.consts = head.mano_dec.mano_layer._consts(self.dev)
        self.cfg = head.loss_cfg
        L = lib()
        R, D, dev = self.R, flow.dim, self.dev
        f = lambda *s: torch.empty(*s, device=dev, dtype=torch.float32)  # noqa: E731
        cpf = L.mhe_flow_cp_floats_per_image(self.shape)
        # static inputs
        # (one contiguous block, every input a 128-byte aligned view of it: a batch collated into staging() travels in ONE host-to-device copy)
        self._input_shapes = {'feat': (B, flow.cond_dim), 'z_det': (B, 16), 'z0': (R, D), 'crop_uv': (B, 42), 'vis': (B, 21)}
        self._input_offsets, n_in = {}, 0
        for k, shp in self._input_shapes.items():
            self._input_offsets[k] = n_in
            n_in += -(-(shp[0] * shp[1]) // 32) * 32
        self.inputs = torch.zeros(n_in, device=dev, dtype=torch.float32)
        self._stage = None
        self.feat, self.z_det, self.z0, self.crop_uv, self.vis = (self._input_view(self.inputs, k) for k in self._input_shapes)
        # forward state
        self.cp, self.x, self.logdet, self.log_q = f(B, cpf), f(R, D), f(R), f(R)
        self.z, self.jtr = f(R, 61), f(R, 21, 3)
        self.verts = f(R, 778, 3) if want_verts else None
        self.uv, self.row_lp = f(R, 42), f(R)
        self.log_p, self.h, self.qlp, self.loss = f(B), f(B), f(B), f(1)
        # gradients
        self.dflat = torch.zeros_like(self.flat)
        self.dcp = f(B, cpf)
        self.djtr, self.dz, self.dlog_q = f(R, 21, 3), f(R, 61), f(R)
        self.dx, self.dz_det, self.dz0, self.dfeat = f(R, D), f(B, 16), f(R, D), f(B, flow.cond_dim)
        self.tc = flow.precision != 'fp32'
        # only the cluster-fused path STORES its weight gradients (mhe_flow_set_async bit 1); the per-GEMM path (long batches, wide
        # masks) accumulates, so the whole gradient buffer is zeroed for it every step
        self.fused = bool(self.tc and L.mhe_flow_pass_is_fused(self.shape, R))
        self.saved = torch.empty(L.mhe_flow_saved_bytes(self.shape, R, int(self.tc)), dtype=torch.uint8, device=dev)
        self.packed = None
        if self.tc:
            # 'optimizer': the flow's own plane set (packed now if stale; FlatAdam.step keeps it current from then on)
            self.packed = flow.packed_weights(dev) if planes == 'optimizer' else \
                torch.empty(L.mhe_flow_packed_bytes(self.shape), dtype=torch.uint8, device=dev)
        self.cws_bytes = L.mhe_flow_cond_workspace_bytes(self.shape, B) if self.tc else 0
        self.cws = torch.empty(max(self.cws_bytes, 16), dtype=torch.uint8, device=dev)
        self.ws_bytes = max(L.mhe_flow_workspace_bytes(self.shape, R, int(self.tc)), L.mhe_mano_workspace_bytes(R, 0))
        self.ws = torch.empty(self.ws_bytes, dtype=torch.uint8, device=dev)
        self.side = torch.cuda.Stream(self.dev)
        self.side2 = torch.cuda.Stream(self.dev)
        self.side3 = torch.cuda.Stream(self.dev)
        self.side4 = torch.cuda.Stream(self.dev)
        self.side5 = torch.cuda.Stream(self.dev)
        self.jtr_mesh = f(R, 21, 3)
        self.mws_bytes = L.mhe_mano_workspace_bytes(R, 0)
        self.mws = torch.empty(self.mws_bytes, dtype=torch.uint8, device=dev)
        self.graph = None
        self.use_graph = use_graph
        # re-plane the weight-gradient operands right after the forward pass (see _enqueue); measured slower, off by default
        self.prepare_ahead = prepare_ahead or _env_on('MHE_ENGINE_PREPARE_AHEAD')
        # mhe_flow_pass_cond_bwd (conditioning backward pipelined into the chunked pass) instead of the two calls: measured equal
        # within 1 % on one GPU (0.557 vs 0.551 ms); it is what a bucketed gradient all-reduce needs (chunk gradients complete early)
        self.pipelined_cond_bwd = pipelined_cond_bwd or _env_on('MHE_ENGINE_PIPELINED_COND_BWD')
        # data parallelism: sum-all-reduce of the flat gradient and the loss INSIDE the step, bucketed by backward
        # chunk - the gradients of a chunk's layers are exchanged while the remaining chunks still run (mhe_flow_join_chunk)
        self.allreduce = bool(allreduce) and torch.distributed.is_available() and torch.distributed.is_initialized() \
            and torch.distributed.get_world_size(allreduce_group) > 1
        self.allreduce_group = allreduce_group
        if self.allreduce:
            self.pipelined_cond_bwd = self.tc
            self.comm = torch.cuda.Stream(self.dev)
        # data parallelism, the default exchange (exchange_gradients(), after the step): the conditioning weight gradient (50 of the 80 MB)
        # has rank <= images per weight matrix, so the ranks all-gather its FACTORS (dcp, feat: 6.4 MB per rank) and each computes the
        # global gradient with mhe_flow_cond_wgrad; only the other 30 MB are all-reduced.  The step itself then skips that GEMM.
        if factored_exchange is None:
            factored_exchange = exchange == 'factored' or (exchange == 'auto' and factored_exchange_pays(self.shape, B, self.world))
        self.factored_exchange = bool(factored_exchange) and self.tc and not self.allreduce and self.world > 1
        # exchange_in_graph (EXPERIMENTAL, off): the factored exchange is enqueued inside the (captured) step, so the gather and the global
        # conditioning GEMM run beside the last weight gradients instead of after the step - 0.638 vs 0.670 ms/step on 2 GPUs in bench.py, but
        # tools/check_factored_exchange.py hung with it (two engines in one process); not the default until that is understood
        self.exchange_in_graph = bool(exchange_in_graph or _env_on('MHE_ENGINE_EXCHANGE_IN_GRAPH')) and self.factored_exchange
        if self.factored_exchange:
            world = self.world
            self.comm = torch.cuda.Stream(self.dev)
            self.comm2 = torch.cuda.Stream(self.dev)
            self.dcp_all = torch.empty(world * B, cpf, device=dev)
            self.feat_all = torch.empty(world * B, flow.cond_dim, device=dev)
            self.xws_bytes = L.mhe_flow_cond_workspace_bytes(self.shape, world * B)
            self.xws = torch.empty(self.xws_bytes, dtype=torch.uint8, device=dev)
        # exchange == 'peer': the gradient and the loss live in symmetric memory and are exchanged INSIDE the step by copy-engine pushes
        # over NVLink (parallel.PeerExchange), bucketed by backward chunk - unlike NCCL's kernels the pushes take no SM from the step's
        # cluster kernels, so chunk c's exchange runs beside the remaining chunks.  Collective: every rank constructs its engine.
        self.px = None
        # MHE_ENGINE_TRACE=1: timestamps of the step's phases (events recorded inside the captured graph; trace_ms() reads them)
        self.trace = [] if _env_on('MHE_ENGINE_TRACE') else None
        if exchange == 'peer' and self.world > 1:
            from .parallel import PeerExchange
            if self.allreduce or self.factored_exchange:
                raise ValueError("exchange='peer' excludes allreduce= / factored_exchange=")
            total = L.mhe_flow_param_floats(self.shape)
            self.px = PeerExchange(total, self.dev, allreduce_group)
            self.px.buf[:total].copy_(self.dflat)
            self.dflat = self.px.buf[:total]
            self.loss = self.px.buf[self.px.n_pad:self.px.n_pad + 1]
            self.pipelined_cond_bwd = self.tc
            self.comm = torch.cuda.Stream(self.dev)
            buckets = [self._chunk_segments(c) for c in range(self._bwd_chunks())]
            buckets[-1].append((self.px.n_pad, self.px.n_pad + 1))       # the loss travels with the last bucket
            self.px.plan(buckets)
            if self.trace is not None:
                self.px.mark = self._mark
        self.launches_per_step = None
        self._g_fwd, self._g_bwd, self.busy = None, {}, False

    def _mark(self, label):
        """Phase mark on the current stream (diagnostic timeline, MHE_ENGINE_TRACE)."""
        if self.trace is None:
            return
        ev = torch.cuda.Event(enable_timing=True, external=True)
        ev.record(torch.cuda.current_stream(self.dev))
        self._trace_now.append((label, ev))

    def trace_ms(self):
        """[(label, ms since the step's start)] of the last run (MHE_ENGINE_TRACE=1)."""
        torch.cuda.synchronize(self.dev)
        t0 = self.trace[0][1]
        return [(label, t0.elapsed_time(ev)) for label, ev in self.trace]

    def _bwd_chunks(self) -> int:
        return lib().mhe_flow_bwd_chunk_count(self.shape, self.R) if self.tc else 1

    def _chunk_segments(self, c: int):
        """Float ranges of the flat gradient that backward chunk ``c`` completes: coupling blocks | conditioning weights | conditioning
        biases of its layers."""
        import ctypes
        L, shape = lib(), self.shape
        total, nlayers = L.mhe_flow_param_floats(shape), shape.layers

        def off(layer, which):      # float offset of (layer, net 0, which); layer == L gives the end of that region
            if layer < nlayers:
                return L.mhe_flow_param_offset(shape, layer, 0, which)
            return {0: L.mhe_flow_param_offset(shape, 0, 0, 6), 6: L.mhe_flow_param_offset(shape, 0, 0, 7), 7: total}[which]

        if not self.tc:
            return [(0, total)]
        l0, nl = ctypes.c_int(), ctypes.c_int()
        check(L.mhe_flow_bwd_chunk_layers(shape, self.R, 0, c, ctypes.byref(l0), ctypes.byref(nl)), 'bwd_chunk_layers')
        return [(off(l0.value, which), off(l0.value + nl.value, which)) for which in (0, 6, 7)]

    def _enqueue_peer_exchange(self, L):
        """Bucketed peer-memory exchange on the communication stream: chunk c's ranges as soon as its layers are complete."""
        main = torch.cuda.current_stream(self.dev)
        nchunks = self._bwd_chunks()
        after = _env_on('MHE_ENGINE_PEER_AFTER')       # diagnostic: the whole exchange after the step's last kernel
        if after or not (self.tc and nchunks > 1):
            self.comm.wait_stream(main)             # no per-chunk events to wait for: everything the main stream enqueued
        with torch.cuda.stream(self.comm):
            sp = _lib.stream_ptr(self.dev)
            for c in range(nchunks):
                if self.tc and nchunks > 1 and not after:
                    check(L.mhe_flow_join_chunk(sp, c), 'join_chunk')
                else:
                    check(L.mhe_flow_join(sp), 'flow_join')
                self._mark(f'chunk{c}.gradients_ready')
                if c == nchunks - 1:
                    self.comm.wait_stream(self.side4)            # the loss (reduced on a side stream)
                self.px.reduce_bucket(c, closing=c == nchunks - 1)
        main.wait_stream(self.comm)

    # ------------------------------------------------------------------
    def _enqueue(self):
        self._trace_now = []
        self._mark('step.start')
        try:
            self._enqueue_body()
        finally:
            self._mark('step.end')
            if self.trace is not None:
                self.trace = self._trace_now

    def _enqueue_body(self):
        L, s = lib(), _lib.stream_ptr(self.dev)
        R, B, shape, ws, wsb = self.R, self.B, self.shape, ptr(self.ws), self.ws_bytes
        z = self.z
        theta, beta = z.data_ptr(), z.data_ptr() + 48 * 4
        pk, cws, cwsb = ptr(self.packed), ptr(self.cws), self.cws_bytes
        # ---- forward
        torch.cuda.nvtx.range_push('mhe.step.prologue')
        repack = self.tc and self.planes == 'step'
        if self.tc:   # weights change between steps.  planes == 'step': refresh their split planes inside the step - only the conditioning
            # planes gate the first GEMM; the coupling planes are converted on a side stream meanwhile, and the bfloat16 copies (read by the
            # backward only) on another one while the forward runs.  planes == 'optimizer': FlatAdam.step wrote them with the update.
            main = torch.cuda.current_stream(self.dev)
            self.side.wait_stream(main)
            if repack and L.mhe_flow_cond_fwd_uses_planes(shape, B):  # (at <= 128 images the conditioning GEMM streams the fp32 weights itself)
                check(L.mhe_flow_pack_weights(shape, ptr(self.flat), pk, 4, s), 'pack_weights')
            if repack:
                self.side3.wait_stream(main)               # after the conditioning planes: those gate the first GEMM
                with torch.cuda.stream(self.side3):
                    check(L.mhe_flow_pack_weights(shape, ptr(self.flat), pk, 8, _lib.stream_ptr(self.dev)), 'pack_weights')
            with torch.cuda.stream(self.side):
                if repack:
                    self.side.wait_stream(self.side3)      # the forward-critical conversions get the memory system first
                    # bfloat16 planes (read by the backward only)
                    check(L.mhe_flow_pack_weights(shape, ptr(self.flat), pk, 2, _lib.stream_ptr(self.dev)), 'pack_weights')
                if self.fused:   # the weight slots of dflat are stored (not accumulated) by the backward: only the bias slots need zeroing
                    check(L.mhe_flow_zero_bias_grads(shape, ptr(self.dflat), _lib.stream_ptr(self.dev)), 'zero_bias_grads')
                else:
                    self.dflat.zero_()
                self.dcp.zero_()
                self.dfeat.zero_()
        check(L.mhe_flow_cond_fwd(shape, ptr(self.flat), pk, ptr(self.feat), B, ptr(self.cp), cws, cwsb, s), 'cond_fwd')
        if repack:
            torch.cuda.current_stream(self.dev).wait_stream(self.side3)
        torch.cuda.nvtx.range_pop()
        torch.cuda.nvtx.range_push('mhe.step.flow_forward')
        check(L.mhe_flow_pass_fwd(shape, ptr(self.flat), pk, ptr(self.mask), ptr(self.cp), ptr(self.z0), R, B, 0, ptr(self.x),
                                  ptr(self.logdet), ptr(self.saved), ws, wsb, s), 'pass_fwd')
        torch.cuda.nvtx.range_pop()
        torch.cuda.nvtx.range_push('mhe.step.hypothesis_rows')
        main = torch.cuda.current_stream(self.dev)
        # (Re-planing the saved activations for the weight gradients right here (mhe_flow_pass_bwd_prepare) and converting the
        # bfloat16 conditioning planes in this window were measured: they delay the per-row kernel and collide with the first
        # backward chunk more than they relieve the forward kernel, so by default the former stays inside mhe_flow_pass_bwd
        # (prepare_ahead = False) and the latter at the start of the step.)
        self.prepared = False
        if self.tc and self.prepare_ahead:
            self.side5.wait_stream(main)
            with torch.cuda.stream(self.side5):
                self.prepared = L.mhe_flow_pass_bwd_prepare(shape, ptr(self.mask), ptr(self.saved), R, 0, ws, wsb,
                                                            _lib.stream_ptr(self.dev)) == 0
        # log q and the image-level reductions are outputs only (the loss is linear in the row terms, so the backward's seeds are
        # constants): they run on a side stream, off the chain flow forward -> z -> per-row kernel -> flow backward
        self.side4.wait_stream(main)
        with torch.cuda.stream(self.side4):
            check(L.mhe_std_normal_logp_fwd(ptr(self.z0), ptr(self.logdet), -1.0, R, shape.dim, ptr(self.log_q),
                                            _lib.stream_ptr(self.dev)), 'log_q')
        check(L.mhe_combine_z_fwd(ptr(self.x), ptr(self.z_det), R, B, ptr(z), s), 'combine_z')
        # The 778-vertex mesh is an output nothing downstream reads: it is skinned on a side stream while the loss and the
        # backward run.
        if self.verts is not None:
            self.side2.wait_stream(main)
            with torch.cuda.stream(self.side2):
                check(L.mhe_mano_fwd(self.consts, theta, 61, beta, 61, R, 1, ptr(self.verts), ptr(self.jtr_mesh), None, ptr(self.mws),
                                     self.mws_bytes, _lib.stream_ptr(self.dev)), 'mano_fwd mesh')
        if self.tc:
            main.wait_stream(self.side)                    # dflat / dcp / dfeat zeroed, bfloat16 planes converted
        else:
            self.dflat.zero_()
            self.dcp.zero_()
        # ---- MANO joints + reprojection / priors forward AND backward (dloss = 1) of every hypothesis: one kernel
        # (the kernel can also assemble its row of z from x / z_det and emit the flow's share of dz itself, which takes both combine_z
        # kernels off this chain - measured slower, 0.549 vs 0.537 ms: the backward's cluster kernel needs entirely free SMs and cannot
        # start before the mesh skinning has drained anyway, and an earlier per-row kernel collides with the pose-blend GEMM)
        check(L.mhe_hypothesis_rows_fwd_bwd(self.consts, self.cfg, ptr(z), None, None, ptr(self.crop_uv), ptr(self.vis), R, B, 1, self.dloss, None,
                                            ptr(self.jtr), ptr(self.uv), ptr(self.row_lp), ptr(self.dz), None, ptr(self.dlog_q), s),
              'hypothesis_rows')
        self.side4.wait_stream(main)
        with torch.cuda.stream(self.side4):
            check(L.mhe_image_loss_reduce(ptr(self.row_lp), ptr(self.log_q), R, B, ptr(self.log_p), ptr(self.h), ptr(self.qlp),
                                          ptr(self.loss), _lib.stream_ptr(self.dev)), 'image_loss_reduce')
            if self.dloss != 1.0:
                self.loss.mul_(self.dloss)              # this rank's share of the global-batch mean (the exchange sums)
        check(L.mhe_combine_z_bwd(ptr(self.dz), R, B, ptr(self.dx), ptr(self.dz_det), s), 'combine_z_bwd')
        torch.cuda.nvtx.range_pop()
        torch.cuda.nvtx.range_push('mhe.step.flow_backward')
        self._mark('flow_backward.start')
        # log_q = log N(z0) - logdet  ->  dL/dlogdet = -dL/dlog_q
        # the weight-gradient GEMMs of the pass keep running on the library's streams while the conditioning backward (which only
        # needs dcp) is enqueued; mhe_flow_join() brings them back before the step ends
        # (bit 1: dflat was zeroed at the start of the step, so the weight-gradient epilogues may store instead of accumulate;
        #  bit 2: so was dfeat)
        if self.tc and self.prepare_ahead:
            torch.cuda.current_stream(self.dev).wait_stream(self.side5)
        flags = ((7 if self.fused else 5) if self.tc else 3) | (8 if self.prepared else 0) | (16 if self.factored_exchange else 0)
        check(L.mhe_flow_set_async(flags), 'set_async')
        if self.px is not None or self.allreduce:
            # the communication stream forks HERE, before the pass is enqueued: it must wait for the chunks' gradient events only
            # (mhe_flow_join_chunk), not for the main stream's position after the whole pass
            self.comm.wait_stream(torch.cuda.current_stream(self.dev))
        try:
            if self.tc and self.pipelined_cond_bwd:
                # ONE call: the conditioning backward is pipelined into the chunked pass
                check(L.mhe_flow_pass_cond_bwd(shape, ptr(self.flat), pk, ptr(self.mask), ptr(self.cp), ptr(self.saved), R, B, 0, ptr(self.dx),
                                               ptr(self.dlog_q), -1.0, ptr(self.dz0), ptr(self.dflat), ptr(self.dcp), ptr(self.feat),
                                               ptr(self.dfeat), ws, wsb, cws, cwsb, s), 'pass_cond_bwd')
            else:
                check(L.mhe_flow_pass_bwd(shape, ptr(self.flat), pk, ptr(self.mask), ptr(self.cp), ptr(self.saved), R, B, 0, ptr(self.dx),
                                          ptr(self.dlog_q), -1.0, ptr(self.dz0), ptr(self.dflat), ptr(self.dcp), ws, wsb, s), 'pass_bwd')
                if self.exchange_in_graph:      # dcp is complete here (stream order): gather the factors, global conditioning GEMM
                    self._enqueue_factor_exchange(L, shape, flags)
                check(L.mhe_flow_cond_bwd(shape, ptr(self.flat), pk, ptr(self.feat), ptr(self.dcp), B, ptr(self.dflat), ptr(self.dfeat),
                                          cws, cwsb, s), 'cond_bwd')
        finally:
            check(L.mhe_flow_set_async(0), 'set_async')
        torch.cuda.nvtx.range_pop()
        torch.cuda.nvtx.range_push('mhe.step.join')
        if self.allreduce:
            self._enqueue_allreduce(L, shape, R)
        self._mark('flow_backward.enqueued')
        if self.px is not None:
            self._enqueue_peer_exchange(L)
        check(L.mhe_flow_join(s), 'flow_join')
        self._mark('flow.joined')
        if self.exchange_in_graph:              # every local gradient is complete: reduce the dense remainder
            self._enqueue_dense_remainder(L, shape)
        if self.verts is not None:
            torch.cuda.current_stream(self.dev).wait_stream(self.side2)    # mesh skinning joins here
        torch.cuda.current_stream(self.dev).wait_stream(self.side4)    # ... and the loss reductions
        torch.cuda.nvtx.range_pop()

    def _enqueue_factor_exchange(self, L, shape, flags):
        """All-gather of the conditioning factors and the global conditioning weight gradient, on the communication stream."""
        import torch.distributed as dist
        grp, main = self.allreduce_group, torch.cuda.current_stream(self.dev)
        self.comm.wait_stream(main)
        with torch.cuda.stream(self.comm):
            # (grouped NCCL launch: each separate call costs ~15-20 us)
            if not _env_on('MHE_ENGINE_NO_COALESCE'):
                with dist._coalescing_manager(group=grp, device=self.dev, async_ops=False):
                    dist.all_gather_into_tensor(self.dcp_all, self.dcp, group=grp)
                    dist.all_gather_into_tensor(self.feat_all, self.feat, group=grp)
            else:
                dist.all_gather_into_tensor(self.dcp_all, self.dcp, group=grp)
                dist.all_gather_into_tensor(self.feat_all, self.feat, group=grp)
            self.gathered = torch.cuda.Event()
            self.gathered.record(self.comm)
        # the GEMM on its own stream (the reductions that follow on the communication stream must not wait for it)
        self.comm2.wait_event(self.gathered)
        with torch.cuda.stream(self.comm2):
            check(L.mhe_flow_set_async(flags), 'set_async')     # (bit 1: the Cw slots are overwritten - they were never written this step)
            check(L.mhe_flow_cond_wgrad(shape, ptr(self.feat_all), ptr(self.dcp_all), self.dcp_all.shape[0], ptr(self.dflat), ptr(self.xws),
                                        self.xws_bytes, _lib.stream_ptr(self.dev)), 'cond_wgrad')

    def _enqueue_dense_remainder(self, L, shape):
        """All-reduce of everything but the conditioning weights (+ the loss) on the communication stream; joins both streams."""
        import torch.distributed as dist
        grp, main = self.allreduce_group, torch.cuda.current_stream(self.dev)
        cw0, cw1 = L.mhe_flow_param_offset(shape, 0, 0, 6), L.mhe_flow_param_offset(shape, 0, 0, 7)
        self.comm.wait_stream(main)
        self.comm.wait_stream(self.side4)           # the loss (reduced on a side stream)
        with torch.cuda.stream(self.comm):
            if not _env_on('MHE_ENGINE_NO_COALESCE'):
                with dist._coalescing_manager(group=grp, device=self.dev, async_ops=False):
                    dist.all_reduce(self.dflat[:cw0], group=grp)
                    dist.all_reduce(self.dflat[cw1:], group=grp)
                    dist.all_reduce(self.loss, group=grp)
            else:
                dist.all_reduce(self.dflat[:cw0], group=grp)
                dist.all_reduce(self.dflat[cw1:], group=grp)
                dist.all_reduce(self.loss, group=grp)
        main.wait_stream(self.comm)
        main.wait_stream(self.comm2)

    def exchange_gradients(self):
        """Data-parallel gradient exchange after run() (sum over the ranks, like one all-reduce of dflat and loss): all-gather of the
        conditioning factors + local mhe_flow_cond_wgrad for the conditioning weights, all-reduce of the rest.  A no-op when the
        exchange already ran inside the step (exchange_in_graph / allreduce)."""
        import torch.distributed as dist
        L, shape, grp = lib(), self.shape, self.allreduce_group
        with _nvtx('mhe.step.exchange_gradients'):
            if not self.factored_exchange:
                if self.world > 1 and not self.allreduce and self.px is None:
                    if _env_on('MHE_ENGINE_NO_COALESCE'):
                        dist.all_reduce(self.dflat, group=grp)
                        dist.all_reduce(self.loss, group=grp)
                    else:       # one grouped NCCL launch for the 80 MB and the scalar
                        with dist._coalescing_manager(group=grp, device=self.dev, async_ops=False):
                            dist.all_reduce(self.dflat, group=grp)
                            dist.all_reduce(self.loss, group=grp)
                return
            if self.exchange_in_graph:
                return
            try:
                self._enqueue_factor_exchange(L, shape, 2 | 16)
            finally:
                check(L.mhe_flow_set_async(0), 'set_async')
            self._enqueue_dense_remainder(L, shape)

    def _enqueue_allreduce(self, L, shape, R):
        """Bucketed sum-all-reduce on the communication stream: chunk c's gradient segments as soon as its layers are complete."""
        import ctypes
        import torch.distributed as dist
        main = torch.cuda.current_stream(self.dev)
        total = L.mhe_flow_param_floats(shape)
        nlayers = shape.layers

        def off(layer, which):      # float offset of (layer, net 0, which); layer == L gives the end of that region
            if layer < nlayers:
                return L.mhe_flow_param_offset(shape, layer, 0, which)
            return {0: L.mhe_flow_param_offset(shape, 0, 0, 6), 6: L.mhe_flow_param_offset(shape, 0, 0, 7), 7: total}[which]

        nchunks = L.mhe_flow_bwd_chunk_count(shape, R) if self.tc else 1
        if not (self.tc and nchunks > 1):
            self.comm.wait_stream(main)
        with torch.cuda.stream(self.comm):
            sp = _lib.stream_ptr(self.dev)
            for c in range(nchunks):
                l0, nl = ctypes.c_int(), ctypes.c_int()
                check(L.mhe_flow_bwd_chunk_layers(shape, R, 0, c, ctypes.byref(l0), ctypes.byref(nl)), 'bwd_chunk_layers')
                if self.tc and nchunks > 1:
                    check(L.mhe_flow_join_chunk(sp, c), 'join_chunk')
                else:
                    check(L.mhe_flow_join(sp), 'flow_join')
                # coupling blocks | conditioning weights | conditioning biases of layers [l0, l0 + nl): one grouped NCCL launch
                segs = [self.dflat[off(l0.value, which):off(l0.value + nl.value, which)] for which in (0, 6, 7)]
                if _env_on('MHE_ENGINE_NO_COALESCE'):
                    for seg in segs:
                        dist.all_reduce(seg, group=self.allreduce_group)
                else:
                    with dist._coalescing_manager(group=self.allreduce_group, device=self.dev, async_ops=False):
                        for seg in segs:
                            dist.all_reduce(seg, group=self.allreduce_group)
            self.comm.wait_stream(self.side4)            # the loss (reduced on a side stream)
            dist.all_reduce(self.loss, group=self.allreduce_group)
        main.wait_stream(self.comm)

    # ------------------------------------------------------------------ split mode: forward and backward as two captured graphs
    # (what the autograd drop-in path replays: MHEntHead.get_loss -> losses._FusedLossFn; the backward receives dL/dlog_p per image)
    def _enqueue_forward_only(self):
        L, s = lib(), _lib.stream_ptr(self.dev)
        R, B, shape, ws, wsb = self.R, self.B, self.shape, ptr(self.ws), self.ws_bytes
        z = self.z
        theta, beta = z.data_ptr(), z.data_ptr() + 48 * 4
        pk, cws, cwsb = ptr(self.packed), ptr(self.cws), self.cws_bytes
        main = torch.cuda.current_stream(self.dev)
        check(L.mhe_flow_cond_fwd(shape, ptr(self.flat), pk, ptr(self.feat), B, ptr(self.cp), cws, cwsb, s), 'cond_fwd')
        check(L.mhe_flow_pass_fwd(shape, ptr(self.flat), pk, ptr(self.mask), ptr(self.cp), ptr(self.z0), R, B, 0, ptr(self.x),
                                  ptr(self.logdet), ptr(self.saved), ws, wsb, s), 'pass_fwd')
        self.side4.wait_stream(main)
        with torch.cuda.stream(self.side4):
            check(L.mhe_std_normal_logp_fwd(ptr(self.z0), ptr(self.logdet), -1.0, R, shape.dim, ptr(self.log_q),
                                            _lib.stream_ptr(self.dev)), 'log_q')
        check(L.mhe_combine_z_fwd(ptr(self.x), ptr(self.z_det), R, B, ptr(z), s), 'combine_z')
        if self.verts is not None:
            self.side2.wait_stream(main)
            with torch.cuda.stream(self.side2):
                check(L.mhe_mano_fwd(self.consts, theta, 61, beta, 61, R, 1, ptr(self.verts), ptr(self.jtr_mesh), None, ptr(self.mws),
                                     self.mws_bytes, _lib.stream_ptr(self.dev)), 'mano_fwd mesh')
        # the per-row kernel's forward outputs (joints, uv, row log-likelihood + priors); its gradient outputs are rewritten by the backward
        check(L.mhe_hypothesis_rows_fwd_bwd(self.consts, self.cfg, ptr(z), None, None, ptr(self.crop_uv), ptr(self.vis), R, B, 1, 0.0, None,
                                            ptr(self.jtr), ptr(self.uv), ptr(self.row_lp), ptr(self.dz), None, None, s), 'hypothesis_rows fwd')
        main.wait_stream(self.side4)
        check(L.mhe_image_loss_reduce(ptr(self.row_lp), ptr(self.log_q), R, B, ptr(self.log_p), ptr(self.h), ptr(self.qlp), ptr(self.loss), s),
              'image_loss_reduce')
        self.norms[0].copy_(z[:, :48].norm(p=2, dim=1))          # th_norm / bt_norm (network.py:787-788): outputs only
        self.norms[1].copy_(z[:, 48:58].norm(p=2, dim=1))
        if self.verts is not None:
            main.wait_stream(self.side2)

    def _enqueue_backward_only(self, dflat):
        """dL/dlog_p per image in self.dlog_p -> dflat (weight slots stored, bias slots zeroed first on the fused path), dfeat, dz_det."""
        L, s = lib(), _lib.stream_ptr(self.dev)
        R, B, shape, ws, wsb = self.R, self.B, self.shape, ptr(self.ws), self.ws_bytes
        pk, cws, cwsb = ptr(self.packed), ptr(self.cws), self.cws_bytes
        main = torch.cuda.current_stream(self.dev)
        self.side.wait_stream(main)
        with torch.cuda.stream(self.side):
            if self.fused:
                check(L.mhe_flow_zero_bias_grads(shape, ptr(dflat), _lib.stream_ptr(self.dev)), 'zero_bias_grads')
            else:
                dflat.zero_()
            self.dcp.zero_()
            self.dfeat.zero_()
        check(L.mhe_hypothesis_rows_fwd_bwd(self.consts, self.cfg, ptr(self.z), None, None, ptr(self.crop_uv), ptr(self.vis), R, B, 1, 1.0,
                                            ptr(self.dlog_p), None, None, ptr(self.row_lp_scratch), ptr(self.dz), None, ptr(self.dlog_q), s),
              'hypothesis_rows bwd')
        check(L.mhe_combine_z_bwd(ptr(self.dz), R, B, ptr(self.dx), ptr(self.dz_det), s), 'combine_z_bwd')
        main.wait_stream(self.side)
        flags = ((7 if self.fused else 5) if self.tc else 3)
        check(L.mhe_flow_set_async(flags), 'set_async')
        try:
            check(L.mhe_flow_pass_bwd(shape, ptr(self.flat), pk, ptr(self.mask), ptr(self.cp), ptr(self.saved), R, B, 0, ptr(self.dx),
                                      ptr(self.dlog_q), -1.0, ptr(self.dz0), ptr(dflat), ptr(self.dcp), ws, wsb, s), 'pass_bwd')
            check(L.mhe_flow_cond_bwd(shape, ptr(self.flat), pk, ptr(self.feat), ptr(self.dcp), B, ptr(dflat), ptr(self.dfeat), cws, cwsb, s),
                  'cond_bwd')
        finally:
            check(L.mhe_flow_set_async(0), 'set_async')
        check(L.mhe_flow_join(s), 'flow_join')

    def _capture(self, fn):
        side = torch.cuda.Stream(self.dev, priority=-1)
        side.wait_stream(torch.cuda.current_stream(self.dev))
        with torch.cuda.stream(side):
            fn()                                  # warm-up outside capture
        torch.cuda.current_stream(self.dev).wait_stream(side)
        torch.cuda.synchronize(self.dev)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            fn()
        return g

    def forward_graph(self):
        if self._g_fwd is None:
            self.norms = torch.empty(2, self.R, device=self.dev)
            self.dlog_p = torch.zeros(self.B, device=self.dev)
            self.row_lp_scratch = torch.empty(self.R, device=self.dev)
            self._g_fwd = self._capture(self._enqueue_forward_only)
        return self._g_fwd

    def backward_graph(self, dflat):
        """Graph of the backward writing its parameter gradients into `dflat` (one graph per destination buffer)."""
        key = dflat.data_ptr()
        if key not in self._g_bwd:
            self._g_bwd[key] = self._capture(lambda: self._enqueue_backward_only(dflat))
        return self._g_bwd[key]

    def _input_view(self, block, k):
        shp, o = self._input_shapes[k], self._input_offsets[k]
        return block[o:o + shp[0] * shp[1]].view(shp)

    def staging(self) -> dict:
        """Pinned host views (feat, z_det, z0, crop_uv, vis) of ONE staging block laid out like the device inputs: a data loader collates
        the batch straight into them and :meth:`load_staged` moves it with a single host-to-device copy."""
        if self._stage is None:
            self._stage_block = torch.zeros(self.inputs.numel(), dtype=torch.float32).pin_memory()
            self._stage = {k: self._input_view(self._stage_block, k) for k in self._input_shapes}
        return self._stage

    def load_staged(self, non_blocking=True):
        """One copy of the whole staging block (see :meth:`staging`) into the static input buffers."""
        self.staging()
        self.inputs.copy_(self._stage_block, non_blocking=non_blocking)

    def load(self, feat, z_det, z0, crop_uv, vis, non_blocking=True):
        """Copy one batch (host or device tensors) into the static input buffers."""
        self.feat.copy_(feat, non_blocking=non_blocking)
        self.z_det.copy_(z_det, non_blocking=non_blocking)
        self.z0.copy_(z0, non_blocking=non_blocking)
        self.crop_uv.copy_(crop_uv, non_blocking=non_blocking)
        self.vis.copy_(vis, non_blocking=non_blocking)

    def run(self):
        """One forward+backward on the loaded batch: ``loss`` (1,), ``dflat``, ``dfeat``, ``dz_det`` are updated."""
        if not self.use_graph:
            n0 = lib().mhe_kernel_launch_count()
            with _nvtx('mhe.step'):
                self._enqueue()
            self.launches_per_step = lib().mhe_kernel_launch_count() - n0
            return self.loss
        if self.graph is None:
            side = torch.cuda.Stream(self.dev, priority=-1)   # the step's critical chain outranks the library's weight-gradient streams
            side.wait_stream(torch.cuda.current_stream(self.dev))
            with torch.cuda.stream(side):
                self._enqueue()                      # warm-up outside capture
            torch.cuda.current_stream(self.dev).wait_stream(side)
            torch.cuda.synchronize(self.dev)
            g = torch.cuda.CUDAGraph()
            n0 = lib().mhe_kernel_launch_count()
            with torch.cuda.graph(g, stream=side):
                self._enqueue()
            self.launches_per_step = lib().mhe_kernel_launch_count() - n0
            self.graph = g
        with _nvtx('mhe.step (graph replay)'):
            self.graph.replay()
        return self.loss

    def flow_grads(self) -> dict:
        """Gradients as reference-named tensors (views into the flat gradient buffer)."""
        flow = self.head.q_z_giv_i
        names = [n for n, _ in flow.named_parameters()]
        by_param = {id(p): n for n, p in flow.named_parameters()}
        return {by_param[id(p)]: self.dflat[off:off + n].view(shape) for p, off, n, shape in flow._slots}
