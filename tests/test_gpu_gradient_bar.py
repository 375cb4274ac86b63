"""The gradient bar, settled with numbers (north star: gradients within 1e-3 relative of the reference).

On the bench shape (BASELINE configs[1], B=64 x S=10) the whole training step - engine and autograd drop-in path - is compared with fp64
autograd of the oracle, separately for the exact-fp32 CUDA path and the tensor-core (split bf16x3) path, next to the reference's OWN
fp32-vs-fp64 distance on the same data (the floor).  See tests/_gradcheck.py for the one structural caveat (leaky-ReLU crossings).
"""
import pytest
import torch

from mhentropy_b200 import MHEntHead
from mhentropy_b200.engine import TrainStep
from mhentropy_b200.mano_assets import synthetic_mano
from mhentropy_b200.synthetic import synthetic_batch
from oracle import flow_oracle as fo, loss_oracle as lo, mano_oracle as mo
from _gradcheck import flat_error, fro, kink_aware_ok, relmax

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def _oracle_step(sd, mano, batch, S, dtype):
    sdg = {k: v.to(dtype).clone().requires_grad_(k != 'mask') for k, v in sd.items()}
    feat = batch['feat'].to(dtype).clone().requires_grad_(True)
    zd = batch['z_det'].to(dtype).clone().requires_grad_(True)
    out = lo.reverse_kld(sdg, mo.mano_constants(mano, dtype), feat, zd, batch['z0'].to(dtype), batch['crop_uv'].to(dtype),
                         batch['vis'].to(dtype), S)
    lo.mhent_loss(out['log_p']).backward()
    return {'log_p': out['log_p'].detach(), 'dfeat': feat.grad, 'dz_det': zd.grad, 'params': {k: v.grad for k, v in sdg.items() if k != 'mask'}}


@pytest.fixture(scope='module')
def bench_case():
    mano, sd = synthetic_mano(0), fo.init_state_dict(seed=0)
    batch = synthetic_batch(64, 10, seed=0)
    ref = _oracle_step(sd, mano, batch, 10, torch.float64)
    o32 = _oracle_step(sd, mano, batch, 10, torch.float32)
    floor = {'dfeat': fro(o32['dfeat'], ref['dfeat']), 'dz_det': fro(o32['dz_det'], ref['dz_det']), 'flat': flat_error(o32['params'], ref['params'])[0]}
    print(f"\nreference fp32 vs fp64 (floor), B=64 x S=10: dfeat {floor['dfeat']:.2e}  dz_det {floor['dz_det']:.2e}  flat {floor['flat']:.2e}")
    return mano, sd, batch, ref, floor


@pytest.mark.parametrize('path', ['engine', 'autograd'])
@pytest.mark.parametrize('precision', ['fp32', 'bf16x3'])
def test_training_step_gradients_meet_1e3_on_the_bench_shape(bench_case, precision, path):
    mano, sd, batch, ref, floor = bench_case
    B, S = 64, 10
    head = MHEntHead(mano_data=mano)
    head.q_z_giv_i.load_state_dict(sd)
    head.q_z_giv_i.precision = precision
    head = head.to(DEV)
    devb = {k: v.to(DEV) for k, v in batch.items()}
    if path == 'engine':
        eng = TrainStep(head, B, S, DEV, want_verts=False, use_graph=True)
        eng.load(**devb)
        eng.run()
        torch.cuda.synchronize()
        got = {'log_p': eng.log_p, 'dfeat': eng.dfeat, 'dz_det': eng.dz_det, 'params': eng.flow_grads()}
    else:
        feat = devb['feat'].clone().requires_grad_(True)
        zd = devb['z_det'].clone().requires_grad_(True)
        for p in head.parameters():
            p.requires_grad_(True)
        out = head.get_loss(feat, {'crop_uv': devb['crop_uv'], 'vis': devb['vis']}, z0=devb['z0'], z_det=zd, N=S)
        (-out['log_p']).mean().backward()
        torch.cuda.synchronize()
        got = {'log_p': out['log_p'], 'dfeat': feat.grad, 'dz_det': zd.grad, 'params': {k: p.grad for k, p in head.q_z_giv_i.named_parameters()}}
    e_lp = relmax(got['log_p'], ref['log_p'])
    e_df, e_dz = fro(got['dfeat'], ref['dfeat']), fro(got['dz_det'], ref['dz_det'])
    e_flat, n_over, worst_name, worst = flat_error(got['params'], ref['params'])
    ok_img, med_img, n_img = kink_aware_ok(got['dfeat'], ref['dfeat'])
    print(f'{path}/{precision} vs fp64: log_p {e_lp:.2e}  dfeat {e_df:.2e} (median image {med_img:.2e}, {n_img} of {B} images above 1e-3)  '
          f'dz_det {e_dz:.2e}  flat gradient {e_flat:.2e} ({n_over} of 240 tensors above 1e-3, worst {worst_name} {worst:.2e})')
    assert e_lp < 1e-4                                    # north star: log_prob within 1e-4 relative
    # north star: gradients within 1e-3 relative - or twice the reference's own fp32 distance to fp64 where that is larger
    assert e_df < max(1e-3, 2 * floor['dfeat']), e_df
    assert e_dz < max(1e-3, 2 * floor['dz_det']), e_dz
    assert e_flat < max(1e-3, 2 * floor['flat']), e_flat
    assert ok_img
    if precision == 'fp32':                               # the exact path carries no scheme error at all
        assert e_df < max(1e-5, 2 * floor['dfeat']) and e_flat < max(1e-5, 2 * floor['flat'])


def test_small_batch_gradients_kink_aware():
    """The smoke shape (B=4 x S=10): with 40 rows a single leaky-ReLU crossing is visible in one image's gradient; the other images meet
    the bar with a wide margin (typ. 1e-5), and the exact-fp32 path meets it outright."""
    mano, sd = synthetic_mano(0), fo.init_state_dict(seed=0)
    B, S = 4, 10
    batch = synthetic_batch(B, S, seed=7)
    ref = _oracle_step(sd, mano, batch, S, torch.float64)
    for precision in ('fp32', 'bf16x3'):
        head = MHEntHead(mano_data=mano)
        head.q_z_giv_i.load_state_dict(sd)
        head.q_z_giv_i.precision = precision
        head = head.to(DEV)
        eng = TrainStep(head, B, S, DEV, want_verts=False, use_graph=False)
        eng.load(**{k: v.to(DEV) for k, v in batch.items()})
        eng.run()
        torch.cuda.synchronize()
        ok, med, n_img = kink_aware_ok(eng.dfeat, ref['dfeat'])
        e = fro(eng.dfeat, ref['dfeat'])
        print(f'B=4 x S=10 {precision}: dfeat {e:.2e}, median image {med:.2e}, {n_img} of {B} images above 1e-3; dz_det {fro(eng.dz_det, ref["dz_det"]):.2e}')
        assert ok and fro(eng.dz_det, ref['dz_det']) < 1e-3
        if precision == 'fp32':
            assert e < 1e-3
