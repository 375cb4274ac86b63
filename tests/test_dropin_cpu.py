"""Host-side checks that need no GPU: the C-ABI library loads and exports every declared symbol, the flat
parameter layout is consistent, and the drop-in classes keep the reference's API / state-dict keys."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from mhentropy_b200 import MHEntHead, ManoLayer, RealNVP, _lib
from mhentropy_b200.build import build
from mhentropy_b200.mano_assets import synthetic_mano
from oracle import flow_oracle as fo

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module', autouse=True)
def _built():
    build()


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, 'include', 'mhentropy_b200.h')).read()
    declared = set(re.findall(r'\b(mhe_[a-z0-9_]+)\s*\(', hdr))
    assert declared == set(_lib.exported_symbols()), declared ^ set(_lib.exported_symbols())
    handle = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(handle, name), name
    assert _lib.lib().mhe_built_for_sm() == 100
    assert _lib.lib().mhe_version() >= 100


def test_flat_layout_matches_reference_parameter_count():
    shape = _lib.FlowShape(45, 512, 512, 12)
    total = _lib.lib().mhe_flow_param_floats(shape)
    flow = RealNVP(dim=45, tsfm_on=512, h_dims=[512, 512], num_steps=6)
    n_params = sum(p.numel() for p in flow.parameters())
    assert n_params == 20_030_520
    assert total >= n_params and total - n_params < 24 * 64 * 2
    # slots are disjoint, aligned and inside the buffer
    spans = []
    for (i, n, which), p in flow._named_flow_params():
        off = _lib.lib().mhe_flow_param_offset(shape, i, n, which)
        assert off % 64 == 0 and off + p.numel() <= total
        spans.append((off, off + p.numel()))
    spans.sort()
    assert all(a[1] <= b[0] for a, b in zip(spans, spans[1:]))
    assert _lib.lib().mhe_flow_cp_floats_per_image(shape) == 12 * 4 * 512


def test_state_dict_keys_match_reference(golden_dir):
    fx = np.load(os.path.join(golden_dir, 'flow_small.npz'))
    ref_keys = {k[2:] for k in fx.files if k.startswith('w/')}
    flow = RealNVP(dim=45, tsfm_on=32, kemb=False, jointN=21, h_dims=[64, 64], num_steps=2)
    assert set(flow.state_dict().keys()) == ref_keys
    flow.load_state_dict({k: torch.from_numpy(fx['w/' + k]) for k in ref_keys})
    head = MHEntHead(q_z_giv_i_cfg=dict(h_dims=[64, 64], num_steps=2, tsfm_on=32), mano_data=synthetic_mano(0), feat_dim=32)
    keys = set(head.state_dict().keys())
    assert 'q_z_giv_i.t.0.l.0.weight' in keys and 'q_z_giv_i.mask' in keys
    for b in ('th_shapedirs', 'th_posedirs', 'th_v_template', 'th_J_regressor', 'th_weights', 'th_faces', 'th_hands_mean',
              'th_comps', 'th_selected_comps', 'th_betas'):
        assert f'mano_dec.mano_layer.{b}' in keys
    assert 'det_head.0.weight' in keys and 'det_head.2.bias' in keys


def test_construction_order_reproduces_reference_weights():
    torch.manual_seed(21)
    flow = RealNVP(dim=45, tsfm_on=16, kemb=False, jointN=21, h_dims=[32, 32], num_steps=3)
    sd = fo.init_state_dict(dim=45, cond_dim=16, h_dims=(32, 32), num_steps=3, seed=21)
    for k, v in flow.state_dict().items():
        assert torch.equal(v, sd[k]), k


def test_cpu_tensors_use_stock_ops_and_match_oracle(golden_dir):
    """API coverage off the kernel path: CPU tensors are evaluated with stock PyTorch ops."""
    fx = np.load(os.path.join(golden_dir, 'flow_small.npz'))
    flow = RealNVP(dim=45, tsfm_on=32, kemb=False, jointN=21, h_dims=[64, 64], num_steps=2)
    flow.load_state_dict({k[2:]: torch.from_numpy(fx[k]) for k in fx.files if k.startswith('w/')})
    x = flow.forward_p(torch.from_numpy(fx['z0']), cond=torch.from_numpy(fx['feat']))
    assert np.abs(x.detach().numpy() - fx['x']).max() < 1e-6
    lp = flow.log_prob(torch.from_numpy(fx['xin']), logvar=torch.from_numpy(fx['feat']))
    assert np.abs(lp.detach().numpy() - fx['log_prob']).max() < 1e-4
    with pytest.raises(NotImplementedError):
        flow.log_prob(torch.from_numpy(fx['xin']), logvar=torch.from_numpy(fx['feat']), weights=torch.zeros(10, 45))


def test_cuda_path_fails_loudly_without_gpu():
    layer = ManoLayer(flat_hand_mean=False, ncomps=45, use_pca=True, skeidx='RHD', mano_data=synthetic_mano(0))
    with pytest.raises(_lib.MheError):
        layer(beta=torch.zeros(2, 10), theta=torch.zeros(2, 48))     # CPU tensors: no fallback for MANO
    with pytest.raises(NotImplementedError):
        ManoLayer(flat_hand_mean=True, ncomps=6, use_pca=True, mano_data=synthetic_mano(0))


def test_invalid_arguments_return_error_codes():
    L = _lib.lib()
    shape = _lib.FlowShape(45, 512, 512, 12)
    st = L.mhe_flow_cond_fwd(shape, None, None, None, 4, None, None, 0, None)
    assert st == 1 and b'null' in L.mhe_last_error_string()
    assert L.mhe_flow_param_offset(_lib.FlowShape(1, 0, 0, 0), 0, 0, 0) == ctypes.c_size_t(-1).value
    assert L.mhe_flow_packed_bytes(shape) > 0 and L.mhe_flow_packed_bytes(_lib.FlowShape(45, 48, 16, 6)) == 0   # hidden % 64
    st = L.mhe_reproj_loss_fwd(ctypes.byref(_lib.LossCfg()), None, None, None, None, None, 7, 2, None, None, None, None, None, None, None)
    assert st == 1


def test_load_reference_checkpoint(tmp_path):
    """A checkpoint in the reference's format (CrossModalHand.py:573-586) loads into MHEntHead without renaming; the backbone entries
    come back untouched; a different flow configuration fails loudly.  Uses the live reference's MHEnt when /root/reference exists."""
    from mhentropy_b200.checkpoint import load_reference_checkpoint
    from oracle import ref_shim
    if ref_shim.reference_available():
        src = ref_shim.build_mhent(synthetic_mano(0), seed=3)
    else:
        torch.manual_seed(3)
        src = MHEntHead(mano_data=synthetic_mano(0))
    sd = {k: v.clone() for k, v in src.state_dict().items()}
    sd['feat_extractor.conv1.weight'] = torch.randn(4, 3, 3, 3)
    path = tmp_path / 'ent_test.pth'
    torch.save({'decoderPose': {}, 'encoderRGB': sd}, path)
    torch.manual_seed(99)
    head = MHEntHead(mano_data=synthetic_mano(1))          # other weights, other MANO constants
    rest = load_reference_checkpoint(head, str(path))
    assert set(rest) == {'feat_extractor.conv1.weight'}
    own = head.state_dict()
    for k, v in sd.items():
        if k in own:
            assert torch.equal(own[k], v), k
    assert len([k for k in own if k.startswith('q_z_giv_i.')]) == 241
    # keep the head's own MANO buffers
    head2 = MHEntHead(mano_data=synthetic_mano(1))
    keep = head2.state_dict()['mano_dec.mano_layer.th_v_template'].clone()
    load_reference_checkpoint(head2, {'encoderRGB': sd}, load_mano_buffers=False)
    assert torch.equal(head2.state_dict()['mano_dec.mano_layer.th_v_template'], keep)
    assert torch.equal(head2.state_dict()['det_head.2.bias'], sd['det_head.2.bias'])
    # a different flow depth must not load silently
    small = MHEntHead(q_z_giv_i_cfg=dict(num_steps=2), mano_data=synthetic_mano(0))
    with pytest.raises(KeyError):
        load_reference_checkpoint(small, {'encoderRGB': sd})


def test_mask_handling_and_fused_limits():
    """A user mask (reference flows.py:131 accepts any) must never run silently wrong: non-{0,1} masks are outside every kernel path,
    and splits wider than the cluster-fused kernels' 24-dim exchange travel in the shape so the ABI routes them to the per-GEMM path."""
    L = _lib.lib()
    flow = RealNVP(dim=45, tsfm_on=512, h_dims=[512, 512], num_steps=6)
    assert flow._kernel_ok and flow._shape.max_split == 23
    m = torch.tensor([[0.] * 15 + [1.] * 30, [1.] * 15 + [0.] * 30] * 6)
    wide = RealNVP(dim=45, tsfm_on=512, h_dims=[512, 512], num_steps=6, mask=m)
    assert wide._kernel_ok and wide._shape.max_split == 30
    # same flat layout / workspace API; the fused path (2 backward chunks) is only offered for the narrow split
    assert L.mhe_flow_param_floats(wide._shape) == L.mhe_flow_param_floats(flow._shape)
    assert L.mhe_flow_bwd_chunk_count(flow._shape, 640) == 2 and L.mhe_flow_bwd_chunk_count(wide._shape, 640) == 1
    assert L.mhe_flow_param_floats(_lib.FlowShape(45, 512, 512, 12, 46)) == 0            # max_split > dim is invalid
    frac = RealNVP(dim=45, tsfm_on=512, h_dims=[512, 512], num_steps=6, mask=m * 0.5)
    assert not frac._kernel_ok and frac._shape is None
    # ... and it still evaluates with the reference's float-mask semantics on CPU tensors
    x = frac.forward_p(torch.randn(3, 45), cond=torch.randn(3, 512))
    assert torch.isfinite(x).all()
    # a checkpoint carrying another mask updates the routing information
    flow.load_state_dict(wide.state_dict())
    assert flow._shape.max_split == 30


def test_mano_assets_are_explicit(tmp_path):
    """No silent synthetic hand: a missing MANO_RIGHT.pkl raises unless a synthetic seed is asked for."""
    with pytest.raises(FileNotFoundError):
        ManoLayer(MANO_dir=str(tmp_path), flat_hand_mean=False, ncomps=45, use_pca=True)
    layer = ManoLayer(MANO_dir=str(tmp_path), flat_hand_mean=False, ncomps=45, use_pca=True, synthetic_seed=0)
    assert layer.mano_layer.mano_source == 'synthetic(seed=0)'
    # derived kernel constants follow a state-dict load (they are keyed on the buffers' versions)
    other = ManoLayer(flat_hand_mean=False, ncomps=45, use_pca=True, mano_data=synthetic_mano(1))
    layer.mano_layer._packed = ('stale',)
    layer.load_state_dict(other.state_dict())
    assert layer.mano_layer._packed is None
    assert torch.equal(layer.mano_layer.th_v_template, other.mano_layer.th_v_template)


def test_reference_mhent_runs_with_the_dropin_flow(golden_dir):
    """INTEGRATION.md section A, executed: the reference's own ``network.MHEnt`` (imported unmodified) with ``RealNVP`` bound to
    ``mhentropy_b200.flows.RealNVP`` reproduces the golden ``get_loss`` outputs and gradients that the unmodified reference produced
    (tests/golden/mhent_small.npz).  CPU tensors: the drop-in's stock-op branch; the kernels' parity is the -m gpu suite."""
    from oracle import ref_shim
    if not ref_shim.reference_available():
        pytest.skip('needs /root/reference (build container only)')
    import sys
    fx = dict(np.load(os.path.join(golden_dir, 'mhent_small.npz')))
    mano = synthetic_mano(0)
    cfg = dict(dim=45, tsfm_on=32, kemb=False, jointN=21, h_dims=[64, 64], num_steps=2)
    network = ref_shim.import_network(mano)
    ref_flows = sys.modules['flows']
    saved = (network.RealNVP, ref_flows.RealNVP)
    network.RealNVP = ref_flows.RealNVP = RealNVP          # the "changed import"
    try:
        model = ref_shim.build_mhent(mano, seed=int(fx['seed']), flow_cfg=cfg)
        assert type(model.q_z_giv_i) is RealNVP
        model.q_z_giv_i.load_state_dict({k[2:]: torch.from_numpy(v) for k, v in fx.items() if k.startswith('w/')})

        class _ZDet(torch.nn.Module):      # stands in for det_head exactly as tests/golden/make_golden.py does
            def __init__(self, z):
                super().__init__()
                self.z = torch.nn.Parameter(z.clone())

            def forward(self, feat):
                return self.z

        model.det_head = _ZDet(torch.from_numpy(fx['z_det']))
        B = fx['feat'].shape[0]
        y = {'crop_uv': torch.from_numpy(fx['crop_uv']), 'vis': torch.from_numpy(fx['vis']), 'st': torch.zeros(B, 3), 'image': np.zeros(1)}
        with ref_shim.cpu_mode():
            feat = torch.from_numpy(fx['feat']).clone().requires_grad_(True)
            torch.manual_seed(int(fx['seed']) + 2)
            out = model.get_loss(feat, y, mods=['uv'])
            (-out['log_p']).mean().backward()
        rel = lambda a, b: float(np.abs(a.detach().numpy() - b).max() / (np.abs(b).max() + 1e-30))  # noqa: E731
        assert rel(out['log_p'], fx['log_p']) < 1e-5
        assert rel(out['h_q_z_giv_i'], fx['h_q_z_giv_i']) < 1e-5
        assert rel(feat.grad, fx['dfeat']) < 1e-4
        assert rel(model.det_head.z.grad, fx['dz_det']) < 1e-4
        for k, p in model.q_z_giv_i.named_parameters():
            assert rel(p.grad, fx['g/' + k]) < 1e-3, k
    finally:
        network.RealNVP, ref_flows.RealNVP = saved
