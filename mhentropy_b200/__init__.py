"""mhentropy_b200 — B200-native multi-hypothesis hot path of GloryyrolG/MHEntropy.

Drop-in classes (same API / state-dict keys as the reference):

* :class:`mhentropy_b200.flows.RealNVP`      <- reference ``hand/flows.py:RealNVP``
* :class:`mhentropy_b200.mano.ManoLayer`     <- reference ``hand/ManoLayer.py:ManoLayer``
* :class:`mhentropy_b200.mano.ManoCore`      <- reference ``hand/manopth/manolayer.py:ManoLayer``
* :class:`mhentropy_b200.losses.MHEntHead`   <- reference ``hand/network.py:MHEnt`` downstream of the feature
* :class:`mhentropy_b200.metrics.MHEntLoss`  <- reference ``hand/criteria.py:MHEntLoss`` (loss + multi-hypothesis metrics)

All CUDA work goes through the C ABI in ``include/mhentropy_b200.h`` (``libmhentropy_b200.so``,
built by ``python -m mhentropy_b200.build``).  There is no CPU fallback for CUDA tensors.
"""
from .flows import RealNVP, _nets  # noqa: F401
from .mano import ManoCore, ManoLayer  # noqa: F401
from .losses import MHEntHead, default_loss_cfg  # noqa: F401
from .metrics import MHEntLoss, hypothesis_metrics, topk_hypotheses  # noqa: F401
from .checkpoint import load_reference_checkpoint  # noqa: F401
from .optim import FlatAdam  # noqa: F401

__all__ = ['RealNVP', 'ManoLayer', 'ManoCore', 'MHEntHead', 'default_loss_cfg', 'MHEntLoss', 'hypothesis_metrics', 'topk_hypotheses', 'load_reference_checkpoint', 'FlatAdam']
