"""The fused engine (mhentropy_b200/engine.py, CUDA-graph replay) against the autograd drop-in path and the oracle."""
import pytest
import torch

from mhentropy_b200 import MHEntHead
from mhentropy_b200.engine import TrainStep
from mhentropy_b200.mano_assets import synthetic_mano
from mhentropy_b200.synthetic import synthetic_batch
from oracle import flow_oracle as fo
from oracle import loss_oracle as lo
from oracle import mano_oracle as mo

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


@pytest.mark.parametrize('use_graph', [False, True])
@pytest.mark.parametrize('B,S', [(8, 10), (5, 3)])
@pytest.mark.parametrize('precision', ['fp32', 'bf16x3'])
def test_engine_matches_autograd_path_and_oracle(use_graph, B, S, precision):
    mano = synthetic_mano(0)
    sd = fo.init_state_dict(seed=0)
    head = MHEntHead(mano_data=mano)
    head.q_z_giv_i.load_state_dict(sd)
    head.q_z_giv_i.precision = precision
    head = head.to(DEV)
    batch = synthetic_batch(B, S, seed=31)
    devb = {k: v.to(DEV) for k, v in batch.items()}
    eng = TrainStep(head, B, S, DEV, want_verts=True, use_graph=use_graph)
    eng.load(**devb)
    for _ in range(2):           # replay twice: the graph must be re-runnable (gradient buffers are re-zeroed inside)
        loss = eng.run()
    torch.cuda.synchronize()
    assert eng.launches_per_step > 10

    feat = devb['feat'].clone().requires_grad_(True)
    z_det = devb['z_det'].clone().requires_grad_(True)
    out = head.get_loss(feat, {'crop_uv': devb['crop_uv'], 'vis': devb['vis']}, z0=devb['z0'], z_det=z_det, N=S)
    loss_ag = (-out['log_p']).mean()
    loss_ag.backward()
    assert rel(loss, loss_ag) < 1e-6
    assert rel(eng.log_p, out['log_p']) < 1e-6
    assert rel(eng.dfeat, feat.grad) < 1e-4
    assert rel(eng.dz_det, z_det.grad) < 1e-4
    g_eng = eng.flow_grads()
    for k, p in head.q_z_giv_i.named_parameters():
        assert rel(g_eng[k], p.grad) < 1e-4, k

    # oracle (fp32, two flow passes like the reference)
    sdg = {k: v.clone().requires_grad_(k != 'mask') for k, v in sd.items()}
    featc = batch['feat'].clone().requires_grad_(True)
    ref = lo.reverse_kld(sdg, mo.mano_constants(mano), featc, batch['z_det'], batch['z0'], batch['crop_uv'], batch['vis'], S)
    lo.mhent_loss(ref['log_p']).backward()
    assert rel(eng.log_p, ref['log_p']) < 1e-4
    # gradient bar by precision: 1e-3 outright on the exact path; kink-aware on the tensor-core path (tests/_gradcheck.py)
    from _gradcheck import kink_aware_ok
    ok, _, n_cross = kink_aware_ok(eng.dfeat, featc.grad)
    assert ok
    assert rel(eng.dfeat, featc.grad) < (1e-3 if (precision == 'fp32' or n_cross == 0) else 5e-3)
    # mesh is materialised in the engine exactly as the reference's get_loss does
    dec = mo.mano_wrapper_forward(mo.mano_constants(mano), ref['z'][:, :48].detach(), ref['z'][:, 48:58].detach())
    assert (eng.verts.cpu() - dec['mesh']).abs().max() < 1e-2


@pytest.mark.parametrize('R,B', [(640, 64), (15, 5), (7, 7)])
def test_hypothesis_rows_kernel_matches_separate_kernels(R, B):
    """mhe_hypothesis_rows_fwd_bwd (one launch) == mano_fwd + reproj_loss_fwd + reproj_loss_bwd + mano_bwd, through the C ABI."""
    from mhentropy_b200 import _lib
    from mhentropy_b200._lib import check, lib, ptr
    head = MHEntHead(mano_data=synthetic_mano(0)).to(DEV)
    consts, cfg, L = head.mano_dec.mano_layer._consts(torch.device(DEV)), head.loss_cfg, lib()
    g = torch.Generator().manual_seed(R)
    z = torch.cat([0.5 * torch.randn(R, 3, generator=g), 1.2 * torch.randn(R, 45, generator=g), 0.03 * torch.randn(R, 10, generator=g),
                   -1.2 + 0.1 * torch.randn(R, 1, generator=g), 0.1 * torch.randn(R, 2, generator=g)], 1).to(DEV).contiguous()
    crop_uv = (torch.rand(B, 42, generator=g) * 2 - 1).to(DEV)
    vis = (torch.rand(B, 21, generator=g) < 0.7).float().to(DEV)
    log_q = torch.randn(R, generator=g).to(DEV)
    s = _lib.stream_ptr(torch.device(DEV))
    f = lambda *sh: torch.empty(*sh, device=DEV)  # noqa: E731
    wsb = L.mhe_mano_workspace_bytes(R, 0)
    ws = torch.empty(wsb, dtype=torch.uint8, device=DEV)
    theta, beta = z.data_ptr(), z.data_ptr() + 48 * 4
    jtr0, uv0, lp0, dj0, dz0, dlq0 = f(R, 21, 3), f(R, 42), f(R), f(R, 21, 3), f(R, 61), f(R)
    logp, h, qlp, loss = f(B), f(B), f(B), f(1)
    check(L.mhe_mano_fwd(consts, theta, 61, beta, 61, R, 1, None, ptr(jtr0), None, ptr(ws), wsb, s), 'mano_fwd')
    check(L.mhe_reproj_loss_fwd(cfg, ptr(jtr0), ptr(z), ptr(crop_uv), ptr(vis), ptr(log_q), R, B, ptr(uv0), ptr(lp0), ptr(logp), ptr(h),
                                ptr(qlp), ptr(loss), s), 'reproj_fwd')
    check(L.mhe_reproj_loss_bwd(cfg, ptr(jtr0), ptr(z), ptr(crop_uv), ptr(vis), R, B, None, None, ptr(dj0), ptr(dz0), ptr(dlq0), s), 'reproj_bwd')
    check(L.mhe_mano_bwd(consts, theta, 61, beta, 61, R, 1, None, ptr(dj0), None, dz0.data_ptr(), 61, dz0.data_ptr() + 48 * 4, 61, 1,
                         ptr(ws), wsb, s), 'mano_bwd')
    jtr1, uv1, lp1, dz1, dlq1 = f(R, 21, 3), f(R, 42), f(R), f(R, 61), f(R)
    check(L.mhe_hypothesis_rows_fwd_bwd(consts, cfg, ptr(z), None, None, ptr(crop_uv), ptr(vis), R, B, 1, 1.0, None, ptr(jtr1), ptr(uv1), ptr(lp1),
                                        ptr(dz1), None, ptr(dlq1), s), 'rows')
    # the same from z's two sources, with the flow's share of dz as an extra output
    x_flow, z_det = z[:, 3:48].contiguous(), torch.cat([z[:B, :3], z[:B, 48:]], 1).contiguous()
    zc = torch.empty_like(z)
    check(L.mhe_combine_z_fwd(ptr(x_flow), ptr(z_det), R, B, ptr(zc), s), 'combine')
    jtr2, uv2, lp2, dz2, dx2 = f(R, 21, 3), f(R, 42), f(R), f(R, 61), f(R, 45)
    check(L.mhe_hypothesis_rows_fwd_bwd(consts, cfg, None, ptr(x_flow), ptr(z_det), ptr(crop_uv), ptr(vis), R, B, 1, 1.0, None, ptr(jtr2), ptr(uv2),
                                        ptr(lp2), ptr(dz2), ptr(dx2), None, s), 'rows from sources')
    jtr3, lp3, dz3 = f(R, 21, 3), f(R), f(R, 61)
    check(L.mhe_hypothesis_rows_fwd_bwd(consts, cfg, ptr(zc), None, None, ptr(crop_uv), ptr(vis), R, B, 1, 1.0, None, ptr(jtr3), None, ptr(lp3),
                                        ptr(dz3), None, None, s), 'rows from combined z')
    torch.cuda.synchronize()
    assert torch.equal(jtr2, jtr3) and torch.equal(lp2, lp3) and torch.equal(dz2, dz3) and torch.equal(dx2, dz2[:, 3:48])
    logp1, loss1 = f(B), f(1)
    check(L.mhe_image_loss_reduce(ptr(lp1), ptr(log_q), R, B, ptr(logp1), None, None, ptr(loss1), s), 'reduce')
    torch.cuda.synchronize()
    assert torch.equal(jtr1, jtr0) and torch.equal(uv1, uv0) and torch.equal(lp1, lp0) and torch.equal(dlq1, dlq0)
    assert rel(dz1, dz0) < 1e-6
    assert torch.equal(logp1, logp) and torch.equal(loss1, loss)


def test_engine_prepare_ahead_matches_default():
    """mhe_flow_pass_bwd_prepare + mhe_flow_set_async bit 3 (weight-gradient operands re-planed right after the forward pass) gives the
    same gradients as the default schedule."""
    head = MHEntHead(mano_data=synthetic_mano(0))
    head.q_z_giv_i.load_state_dict(fo.init_state_dict(seed=0))
    head.q_z_giv_i.precision = 'bf16x3'
    head = head.to(DEV)
    devb = {k: v.to(DEV) for k, v in synthetic_batch(64, 10, seed=5).items()}
    outs = []
    for ahead in (False, True):
        eng = TrainStep(head, 64, 10, DEV, want_verts=False, use_graph=True, prepare_ahead=ahead)
        eng.load(**devb)
        for _ in range(2):
            eng.run()
        torch.cuda.synchronize()
        outs.append((eng.loss.clone(), eng.dflat.clone(), eng.dfeat.clone(), eng.dz0.clone()))
    assert torch.equal(outs[0][0], outs[1][0])
    for a, b in zip(outs[0][1:], outs[1][1:]):
        assert rel(a, b) < 1e-5


def test_engine_pipelined_cond_bwd_matches_default():
    """mhe_flow_pass_cond_bwd (one call, conditioning backward pipelined into the chunked pass) == mhe_flow_pass_bwd + mhe_flow_cond_bwd."""
    head = MHEntHead(mano_data=synthetic_mano(0))
    head.q_z_giv_i.load_state_dict(fo.init_state_dict(seed=0))
    head.q_z_giv_i.precision = 'bf16x3'
    head = head.to(DEV)
    for B, S in ((64, 10), (5, 3)):
        devb = {k: v.to(DEV) for k, v in synthetic_batch(B, S, seed=6).items()}
        outs = []
        for piped in (False, True):
            eng = TrainStep(head, B, S, DEV, want_verts=False, use_graph=True, pipelined_cond_bwd=piped)
            eng.load(**devb)
            for _ in range(2):
                eng.run()
            torch.cuda.synchronize()
            outs.append((eng.loss.clone(), eng.dflat.clone(), eng.dfeat.clone(), eng.dz0.clone()))
        assert torch.equal(outs[0][0], outs[1][0])
        for a, b in zip(outs[0][1:], outs[1][1:]):
            assert rel(a, b) < 1e-5


def test_cond_wgrad_matches_cond_bwd():
    """mhe_flow_cond_wgrad (the conditioning weight gradient from its factors, used by the factored data-parallel exchange) writes exactly
    what mhe_flow_cond_bwd writes into the Cw slots; with mhe_flow_set_async bit 4 cond_bwd leaves those slots alone."""
    from mhentropy_b200 import RealNVP, _lib
    from mhentropy_b200._lib import check, lib, ptr
    torch.manual_seed(3)
    flow = RealNVP(dim=45, tsfm_on=512, h_dims=[512, 512], num_steps=6).to(DEV)
    flow.precision = 'bf16x3'
    L, shape = lib(), flow._shape
    flat, packed = flow.flat_parameters(torch.device(DEV)), flow.packed_weights(torch.device(DEV))
    s = _lib.stream_ptr(torch.device(DEV))
    cw0, cw1 = L.mhe_flow_param_offset(shape, 0, 0, 6), L.mhe_flow_param_offset(shape, 0, 0, 7)
    for B in (64, 5):
        feat = torch.randn(B, 512, device=DEV)
        dcp = 1e-3 * torch.randn(B, L.mhe_flow_cp_floats_per_image(shape), device=DEV)
        wsb = L.mhe_flow_cond_workspace_bytes(shape, B)
        ws = torch.empty(wsb, dtype=torch.uint8, device=DEV)
        ga, gb = torch.zeros_like(flat), torch.zeros_like(flat)
        dfa, dfb = torch.empty(B, 512, device=DEV), torch.empty(B, 512, device=DEV)
        check(L.mhe_flow_cond_bwd(shape, ptr(flat), ptr(packed), ptr(feat), ptr(dcp), B, ptr(ga), ptr(dfa), ptr(ws), wsb, s), 'cond_bwd')
        check(L.mhe_flow_set_async(16), 'set_async')
        try:
            check(L.mhe_flow_cond_bwd(shape, ptr(flat), ptr(packed), ptr(feat), ptr(dcp), B, ptr(gb), ptr(dfb), ptr(ws), wsb, s), 'cond_bwd skip')
        finally:
            check(L.mhe_flow_set_async(0), 'set_async')
        torch.cuda.synchronize()
        assert float(gb[cw0:cw1].abs().max()) == 0.0                      # Cw slots untouched
        assert rel(gb[:cw0], ga[:cw0]) < 1e-5 and rel(gb[cw1:], ga[cw1:]) < 1e-5        # biases as before (atomic sums: order may differ)
        assert rel(dfb, dfa) < 1e-5
        check(L.mhe_flow_cond_wgrad(shape, ptr(feat), ptr(dcp), B, ptr(gb), ptr(ws), wsb, s), 'cond_wgrad')
        torch.cuda.synchronize()
        assert torch.equal(gb[cw0:cw1], ga[cw0:cw1])
        # and against plain PyTorch (fp64)
        from mhentropy_b200.parallel import cond_wgrad_from_factors
        ref = cond_wgrad_from_factors(dcp.double().cpu(), feat.double().cpu(), 512)
        got = torch.stack([ga[L.mhe_flow_param_offset(shape, i // 4, (i // 2) % 2, 6 + 2 * (i % 2)):][:512 * 512].view(512, 512) for i in range(48)])
        assert rel(got, ref) < 1e-3      # gradients travel as bfloat16 split planes (16 significant bits): the north-star gradient bar


def test_flat_adam_matches_torch_adam_and_refreshes_planes():
    """mhe_flow_adam_step == torch.optim.Adam on the flat buffer (reference CrossModalHand.py:201, clip scale :462-470), and the split
    planes it leaves behind are byte-identical to a fresh mhe_flow_pack_weights of the updated parameters."""
    from mhentropy_b200 import RealNVP, _lib
    from mhentropy_b200._lib import check, lib, ptr
    from mhentropy_b200.optim import FlatAdam
    torch.manual_seed(8)
    dev = torch.device(DEV)
    flow = RealNVP(dim=45, tsfm_on=512, h_dims=[512, 512], num_steps=6).to(dev)
    flow.precision = 'bf16x3'
    L, shape = lib(), flow._shape
    flat = flow.flat_parameters(dev)
    packed = flow.packed_weights(dev)
    ref_p = torch.nn.Parameter(flat.detach().clone())
    ref_opt = torch.optim.Adam([ref_p], lr=2e-3)
    opt = FlatAdam(flow, lr=2e-3)
    for it in range(3):
        g = torch.zeros_like(flat)                                   # like a real flat gradient: zero on the padding between tensors
        for _, off, n, _ in flow._slots:
            g[off:off + n] = 1e-2 * torch.randn(n, device=dev)
        sq = opt.grad_sqnorm(g)
        assert abs(float(sq) - float(g.double().pow(2).sum())) < 1e-9 * float(sq)
        scale = min(1.0, 1.0 / (float(sq) ** 0.5 + 1e-6))             # clip_grad_norm_(…, 1.)
        ref_p.grad = g * scale
        ref_opt.step()
        opt.step(g, grad_scale=scale)
    torch.cuda.synchronize()
    err = rel(flow.flat_parameters(dev), ref_p)
    assert err < 2e-6, err
    fresh = packed.clone()             # (same padding bytes between the plane families; the data regions are re-packed below)
    check(L.mhe_flow_pack_weights(shape, ptr(flow.flat_parameters(dev)), ptr(fresh), 3, _lib.stream_ptr(dev)), 'pack_weights')
    torch.cuda.synchronize()
    # compare the plane regions only (padding between them is never written)
    nb = L.mhe_flow_packed_bytes(shape)
    a, b = packed[:nb].view(torch.int16), fresh[:nb].view(torch.int16)
    assert int((a != b).sum()) == 0


def test_data_parallel_engine_equals_single_process_step():
    """TrainStep + exchange_gradients on 2 GPUs == one TrainStep on the concatenated batch (global-batch-mean loss and gradients).
    Needs two GPUs on the box (the driver's GPU tier has one: it is run with `gpurun --gpus 2`, log under profiles/)."""
    import os
    import subprocess
    import sys
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr', '127.0.0.1',
                        '--master-port', '29533', os.path.join(root, 'tools', 'check_engine_dp.py')], capture_output=True, text=True, timeout=600)
    print(r.stdout[-2000:], r.stderr[-2000:])
    assert r.returncode == 0 and 'check_engine_dp: OK' in r.stdout


def test_engine_long_batch_replay_matches_autograd_path():
    """Above the cluster-fused row limit the engine runs the per-GEMM path, whose weight gradients ACCUMULATE (split contractions): the
    step must zero the whole gradient buffer itself, or a replayed step would add to the previous one (found in round 2)."""
    head = MHEntHead(mano_data=synthetic_mano(0))
    head.q_z_giv_i.load_state_dict(fo.init_state_dict(seed=0))
    head.q_z_giv_i.precision = 'bf16x3'
    head = head.to(DEV)
    B, S = 66, 65
    devb = {k: v.to(DEV) for k, v in synthetic_batch(B, S, seed=17).items()}
    eng = TrainStep(head, B, S, DEV, want_verts=False, use_graph=True)
    assert not eng.fused
    eng.load(**devb)
    for _ in range(3):
        loss = eng.run()
    torch.cuda.synchronize()
    feat = devb['feat'].clone().requires_grad_(True)
    z_det = devb['z_det'].clone().requires_grad_(True)
    for p in head.parameters():
        p.requires_grad_(True)
    head.zero_grad(set_to_none=True)
    out = head.get_loss(feat, {'crop_uv': devb['crop_uv'], 'vis': devb['vis']}, z0=devb['z0'], z_det=z_det, N=S)
    (-out['log_p']).mean().backward()
    assert rel(loss, (-out['log_p']).mean()) < 1e-6
    assert rel(eng.dfeat, feat.grad) < 1e-4 and rel(eng.dz_det, z_det.grad) < 1e-4
    g_eng = eng.flow_grads()
    for k, p in head.q_z_giv_i.named_parameters():
        assert rel(g_eng[k], p.grad) < 1e-4, k


def test_fused_loss_node_general_downstream_and_accumulation():
    """MHEntHead.get_loss through the fused autograd node (two captured graphs) against the unfused chain of autograd functions: a
    NON-uniform downstream loss (weights per image), gradient accumulation over two backward calls, and a dropped graph."""
    head = MHEntHead(mano_data=synthetic_mano(0))
    head.q_z_giv_i.load_state_dict(fo.init_state_dict(seed=0))
    head.q_z_giv_i.precision = 'bf16x3'
    head = head.to(DEV)
    for p in head.parameters():
        p.requires_grad_(True)
    B, S = 6, 10
    devb = {k: v.to(DEV) for k, v in synthetic_batch(B, S, seed=41).items()}
    w = torch.linspace(0.5, 2.0, B, device=DEV)

    def run(fuse, feat, zd):
        head.fuse_loss = fuse
        out = head.get_loss(feat, {'crop_uv': devb['crop_uv'], 'vis': devb['vis']}, z0=devb['z0'], z_det=zd, N=S)
        return out, -(w * out['log_p']).sum()

    res = {}
    for fuse in (False, True):
        head.zero_grad(set_to_none=True)
        feat, zd = devb['feat'].clone().requires_grad_(True), devb['z_det'].clone().requires_grad_(True)
        if fuse:
            run(True, feat, zd)                     # a forward whose graph is dropped without backward must release its engine
        out, loss = run(fuse, feat, zd)
        loss.backward()
        out2, loss2 = run(fuse, feat, zd)           # second backward: gradients accumulate (2x)
        loss2.backward()
        torch.cuda.synchronize()
        res[fuse] = (out, feat.grad.clone(), zd.grad.clone(), {k: p.grad.clone() for k, p in head.q_z_giv_i.named_parameters()})
    (o0, f0, z0g, g0), (o1, f1, z1g, g1) = res[False], res[True]
    for k in ('log_p', 'h_q_z_giv_i', 'q_log_p_z_giv_y', 'uv_mu', 'th_norm', 'bt_norm'):
        assert rel(o1[k], o0[k]) < 1e-5, k
    assert rel(f1, f0) < 1e-4 and rel(z1g, z0g) < 1e-4
    for k in g0:
        assert rel(g1[k], g0[k]) < 1e-4, k
    assert all(not e.busy for pool in head._fused_pool.values() for e in pool)


def test_staged_single_copy_load_equals_per_tensor_load():
    """TrainStep.staging() + load_staged() (one host-to-device copy of the pinned block) gives the same step as load() of five tensors; an
    odd shape exercises the padded segment offsets."""
    head = MHEntHead(mano_data=synthetic_mano(0))
    head.q_z_giv_i.load_state_dict(fo.init_state_dict(seed=0))
    head.q_z_giv_i.precision = 'bf16x3'
    head = head.to(DEV)
    B, S = 7, 9
    batch = synthetic_batch(B, S, seed=23)
    eng = TrainStep(head, B, S, DEV, want_verts=False, use_graph=True)
    eng.load(**{k: v.to(DEV) for k, v in batch.items()})
    l0 = eng.run().clone()
    g0, f0 = eng.dflat.clone(), eng.dfeat.clone()
    for v in eng.staging().values():
        assert v.is_pinned()
    eng.inputs.zero_()
    for k, v in batch.items():
        eng.staging()[k].copy_(v)
    eng.load_staged()
    l1 = eng.run().clone()
    torch.cuda.synchronize()
    assert rel(l1, l0) < 1e-6 and rel(eng.dflat, g0) < 1e-5 and rel(eng.dfeat, f0) < 1e-5      # (atomic bias sums: not bit-stable)
