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
    assert rel(eng.dfeat, featc.grad) < 5e-3
    # mesh is materialised in the engine exactly as the reference's get_loss does
    dec = mo.mano_wrapper_forward(mo.mano_constants(mano), ref['z'][:, :48].detach(), ref['z'][:, 48:58].detach())
    assert (eng.verts.cpu() - dec['mesh']).abs().max() < 1e-2
