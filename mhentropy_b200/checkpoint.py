"""Loading the reference's checkpoints into the drop-in modules.

The reference saves ``{'decoderPose': ..., 'encoderRGB': ..., ['p_nf': ...]}`` (``hand/CrossModalHand.py:573-586``) and loads it module
by module (``:588-602``).  ``encoderRGB`` is the ``MHEnt`` network (``network.py:309``): its hot-path sub-modules keep the reference's
parameter names here (``q_z_giv_i.{s,t}.{i}.{l,c}.{j}.{weight,bias}``, ``q_z_giv_i.mask``, ``mano_dec.mano_layer.th_*``,
``det_head.{0,2}.*``), so the released ``model/ent_ho3d.pth`` (``ho3d.yaml:54``) loads without renaming; everything else in it
(``feat_extractor.*``: the CNN backbone, outside the measured path) is handed back for the caller's stock PyTorch backbone.
"""
from __future__ import annotations

import torch

HOT_PATH_PREFIXES = ('q_z_giv_i.', 'mano_dec.', 'det_head.')


def split_reference_state_dict(state_dict: dict) -> tuple[dict, dict]:
    """(hot-path entries, the rest) of an ``MHEnt`` state dict."""
    hot = {k: v for k, v in state_dict.items() if k.startswith(HOT_PATH_PREFIXES)}
    rest = {k: v for k, v in state_dict.items() if not k.startswith(HOT_PATH_PREFIXES)}
    return hot, rest


def load_reference_checkpoint(head, checkpoint, module: str = 'encoderRGB', load_mano_buffers: bool = True) -> dict:
    """Load the hot-path part of a reference checkpoint into ``head`` (:class:`mhentropy_b200.losses.MHEntHead`).

    checkpoint: a path (``torch.load``-ed on the CPU), the checkpoint dict the reference saves, or an ``MHEnt`` state dict itself.
    load_mano_buffers: also take the MANO buffers stored in the checkpoint (they are the constants of ``MANO_RIGHT.pkl``); set False
    to keep the ones ``head`` was built with.  Returns the entries that are not on the hot path (the backbone), untouched.
    Raises ``KeyError`` when flow / ``det_head`` parameters are missing or unexpected (same strictness as the reference's
    ``load_state_dict``), so a checkpoint of a different flow configuration fails loudly.
    """
    if isinstance(checkpoint, (str, bytes)) or hasattr(checkpoint, '__fspath__'):
        checkpoint = torch.load(checkpoint, map_location='cpu', weights_only=True)   # state dicts are plain tensors: no unpickling of code
    sd = checkpoint[module] if module in checkpoint and isinstance(checkpoint[module], dict) else checkpoint
    hot, rest = split_reference_state_dict(sd)
    if not load_mano_buffers:
        hot = {k: v for k, v in hot.items() if not k.startswith('mano_dec.')}
    own = head.state_dict()
    wanted = {k for k in own if load_mano_buffers or not k.startswith('mano_dec.')}
    missing, unexpected = sorted(wanted - set(hot)), sorted(set(hot) - set(own))
    if missing or unexpected:
        raise KeyError(f'reference checkpoint does not match the head: missing {missing[:5]}{"..." if len(missing) > 5 else ""}, '
                       f'unexpected {unexpected[:5]}{"..." if len(unexpected) > 5 else ""}')
    head.load_state_dict(hot, strict=load_mano_buffers)
    # derived state follows the loaded tensors: the flow's split weight planes and the MANO kernels' packed constants are rebuilt on
    # next use (both modules also invalidate themselves in _load_from_state_dict; an engine built BEFORE the load must be rebuilt -
    # TrainStep bakes the constants' pointers into its CUDA graph)
    flow = getattr(head, 'q_z_giv_i', None)
    if flow is not None and hasattr(flow, 'mark_parameters_changed'):
        flow.mark_parameters_changed()
    mano = getattr(getattr(head, 'mano_dec', None), 'mano_layer', None)
    if mano is not None and hasattr(mano, '_packed'):
        mano._packed = None
    return rest
