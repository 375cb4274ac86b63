"""Oracle support (TEST INFRASTRUCTURE): import the UNMODIFIED reference in the build container.

Only usable where ``/root/reference`` is mounted (never on the GPU box).  Follows the recipe of
SURVEY.md §8c: stub the absent third-party modules *outside* the reference tree, hand the
reference synthetic MANO-shaped constants through a fake ``ready_arguments`` and neutralise its
hard-coded ``.cuda()`` / ``device='cuda'`` sites with a ``TorchFunctionMode`` so it runs on CPU.
Nothing from the reference is copied; it is imported from where it lies.
"""
from __future__ import annotations

import contextlib
import os
import sys
import types

import numpy as np
import torch
from torch.overrides import TorchFunctionMode

REFERENCE_ROOT = os.environ.get('MHE_REFERENCE_ROOT', '/root/reference')
REFERENCE_HAND = os.path.join(REFERENCE_ROOT, 'hand')


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_HAND, 'flows.py'))


class _R:
    """chumpy stand-in: the reference only ever reads ``.r`` (``manolayer.py:71-82``)."""

    def __init__(self, a):
        self.r = np.asarray(a)


class _Sparse:
    """scipy-sparse stand-in: the reference only calls ``.toarray()`` (``manolayer.py:79``)."""

    def __init__(self, a):
        self._a = np.asarray(a)

    def toarray(self):
        return self._a


_MANO_DICT = None


def _ready_arguments(_path):
    m = _MANO_DICT
    return {
        'betas': _R(m['betas']), 'shapedirs': _R(m['shapedirs']), 'posedirs': _R(m['posedirs']),
        'v_template': _R(m['v_template']), 'J_regressor': _Sparse(m['J_regressor']),
        'weights': _R(m['weights']), 'f': m['f'], 'hands_components': np.asarray(m['hands_components']),
        'hands_mean': np.asarray(m['hands_mean']), 'kintree_table': np.asarray(m['kintree_table']),
    }


def _stub(name, **attrs):
    mod = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(mod, k, v)
    sys.modules[name] = mod
    return mod


def install_stubs(mano: dict):
    """Put stub modules in ``sys.modules`` before the reference is imported."""
    global _MANO_DICT
    _MANO_DICT = mano
    if 'mano.webuser.smpl_handpca_wrapper_HAND_only' not in sys.modules:
        _stub('mano')
        _stub('mano.webuser')
        _stub('mano.webuser.smpl_handpca_wrapper_HAND_only', ready_arguments=_ready_arguments)
    for name in ('pycocotools', 'trimesh', 'matplotlib', 'matplotlib.pyplot', 'matplotlib.animation',
                 'mpl_toolkits', 'mpl_toolkits.mplot3d'):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                _stub(name)
    if 'pycocotools.coco' not in sys.modules:
        _stub('pycocotools.coco', COCO=object)
        _stub('pycocotools.cocoeval', COCOeval=object)
    if 'nflows' not in sys.modules:
        _stub('nflows')
        _stub('nflows.flows', ConditionalGlow=type('ConditionalGlow', (torch.nn.Module,), {}))
    if REFERENCE_HAND not in sys.path:
        sys.path.insert(0, REFERENCE_HAND)


class CudaToCpu(TorchFunctionMode):
    """Rewrite ``device='cuda'`` -> 'cpu' and make ``Tensor.cuda()`` a no-op."""

    def __torch_function__(self, func, types_, args=(), kwargs=None):
        kwargs = dict(kwargs or {})
        dev = kwargs.get('device')
        if dev is not None and 'cuda' in str(dev):
            kwargs['device'] = 'cpu'
        if func is torch.Tensor.cuda:
            return args[0]
        if func is torch.Tensor.to and len(args) > 1 and isinstance(args[1], (str, torch.device)) and 'cuda' in str(args[1]):
            args = (args[0], 'cpu') + tuple(args[2:])
        return func(*args, **kwargs)


@contextlib.contextmanager
def cpu_mode():
    if torch.cuda.is_available():
        yield
    else:
        with CudaToCpu():
            yield


def import_flows(mano: dict):
    install_stubs(mano)
    import flows  # reference hand/flows.py
    return flows


def import_manolayer(mano: dict):
    install_stubs(mano)
    with cpu_mode():
        import ManoLayer as ref_manolayer  # reference hand/ManoLayer.py
    return ref_manolayer


def import_network(mano: dict):
    install_stubs(mano)
    with cpu_mode():
        import network  # reference hand/network.py
    return network


def build_mhent(mano: dict, seed: int = 0, flow_cfg: dict | None = None):
    """Construct the reference ``MHEnt`` (``network.py:309``) with the HO3D configuration and a
    pass-through feature extractor: ``x`` *is* the 512-d feature."""
    network = import_network(mano)
    cfg = dict(dim=45, tsfm_on=512, kemb=False, jointN=21, h_dims=[512, 512], num_steps=6)
    if flow_cfg:
        cfg.update(flow_cfg)

    class _PassThrough(network.BasicEnc):
        def __init__(self):
            torch.nn.Module.__init__(self)

        def forward(self, x, deterministic=False, p=None):
            return None, x, None

    special = {
        'q_z_giv_i_model': 'realnvp', 'q_z_giv_i_cfg': cfg, 'ds': 'ho3d', 'image_size': [256, 256],
        'mano_cfg': {'flat_hand_mean': False, 'ncomps': 45, 'use_pca': True},
        'prior_cfg': {'p_theta45_pth': None, 'th45_ref_alpha': 50},
        'data_prior_cfg': {'b_2d': 0.03, 'w_prior_2d': 0},
        'loss_cfg': {'entropy': True, 'mode': False, 'w_reg_ds': 0},
        'kld_w': 1, 'kld_w_annealing': [1, 20 * 1200], 'T': 1.0,
    }
    common = dict(n_latent=512, backbone='resnet18', pretrained=False, conditional_p=False, K=21, D=3,
                  feat_dim=None, sigma_act='exp', deterministic=False, input='image')
    with cpu_mode():
        torch.manual_seed(seed)
        # the flow is the first randomly initialised sub-module we care about: build it under the
        # seed first so its weights equal flow_oracle.init_state_dict(seed)
        flows = sys.modules['flows']
        flow = flows.RealNVP(**cfg)
        model = network.MHEnt(special, **common)
        model.q_z_giv_i = flow
        model.feat_extractor = _PassThrough()
    return model
