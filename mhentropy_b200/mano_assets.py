"""MANO model constants: loader for a real ``MANO_RIGHT.pkl`` and a seeded synthetic stand-in.

The reference reads its constants through ``mano.webuser...ready_arguments`` and registers them
as buffers (reference ``hand/manopth/manolayer.py:61-108``).  ``MANO_RIGHT.pkl`` is licence-gated
and absent from this build environment, so benchmarks and parity tests use
:func:`synthetic_mano` — MANO-*shaped* arrays drawn from a fixed ``numpy.random.RandomState`` —
exactly as SURVEY.md §8c.3 prescribes.  The same dictionary feeds the reference (through the
shim in ``oracle/ref_shim.py``), the oracle and the CUDA path, so all three see identical
constants.

Keys (all ``numpy`` arrays, float64 unless noted), mirroring what ``manolayer.py`` consumes:

``betas (10,)``, ``shapedirs (778,3,10)``, ``posedirs (778,3,135)``, ``v_template (778,3)``,
``J_regressor (16,778)``, ``weights (778,16)``, ``f (1538,3) uint32``,
``hands_components (45,45)``, ``hands_mean (45,)``, ``kintree_table (2,16) int64``.
"""
from __future__ import annotations

import os
import pickle

import numpy as np

N_VERTS = 778
N_JOINTS = 16
N_FACES = 1538
N_POSE = 45
N_SHAPE = 10

# parents of the 16 MANO joints (reference: smpl_data['kintree_table'][0], manolayer.py:104-108).
KINTREE_PARENTS = [-1, 0, 1, 2, 0, 4, 5, 0, 7, 8, 0, 10, 11, 0, 13, 14]


def synthetic_mano(seed: int = 0) -> dict:
    """MANO-shaped constants with plausible magnitudes (metres / radians), reproducible from ``seed``."""
    rs = np.random.RandomState(seed)
    v_template = rs.uniform(-0.09, 0.09, size=(N_VERTS, 3))
    shapedirs = rs.normal(0.0, 0.01, size=(N_VERTS, 3, N_SHAPE))
    posedirs = rs.normal(0.0, 0.004, size=(N_VERTS, 3, 9 * 15))
    # joint regressor: each joint is a convex combination of a handful of vertices
    J_regressor = np.zeros((N_JOINTS, N_VERTS))
    for j in range(N_JOINTS):
        idx = rs.choice(N_VERTS, size=24, replace=False)
        w = rs.uniform(0.1, 1.0, size=24)
        J_regressor[j, idx] = w / w.sum()
    # skinning weights: each vertex follows up to four joints, rows sum to one
    weights = np.zeros((N_VERTS, N_JOINTS))
    for v in range(N_VERTS):
        idx = rs.choice(N_JOINTS, size=4, replace=False)
        w = rs.uniform(0.0, 1.0, size=4) ** 2
        weights[v, idx] = w / w.sum()
    faces = rs.randint(0, N_VERTS, size=(N_FACES, 3)).astype(np.uint32)
    q, _ = np.linalg.qr(rs.normal(size=(N_POSE, N_POSE)))
    hands_components = q * rs.uniform(0.2, 1.2, size=(N_POSE, 1))
    hands_mean = rs.normal(0.0, 0.15, size=(N_POSE,))
    betas = np.zeros((N_SHAPE,))
    kintree = np.array([[4294967295] + KINTREE_PARENTS[1:], list(range(N_JOINTS))], dtype=np.int64)
    return {
        'betas': betas,
        'shapedirs': shapedirs,
        'posedirs': posedirs,
        'v_template': v_template,
        'J_regressor': J_regressor,
        'weights': weights,
        'f': faces,
        'hands_components': hands_components,
        'hands_mean': hands_mean,
        'kintree_table': kintree,
    }


def _dense(a) -> np.ndarray:
    """chumpy objects expose ``.r``; scipy sparse matrices ``.toarray()``; arrays pass through."""
    if hasattr(a, 'toarray'):
        a = a.toarray()
    if hasattr(a, 'r'):
        a = a.r
    return np.asarray(a)


def load_mano_pkl(path: str) -> dict:
    """Read a real ``MANO_RIGHT.pkl`` into the same dictionary :func:`synthetic_mano` returns.

    Only needs ``pickle`` + ``numpy`` when the pickle was re-exported without chumpy; with the
    original SMPL+H pickle, ``chumpy`` and ``scipy`` must be importable for unpickling.
    """
    with open(path, 'rb') as fh:
        raw = pickle.load(fh, encoding='latin1')
    posedirs = _dense(raw['posedirs'])
    out = {
        'betas': np.zeros((N_SHAPE,)) if 'betas' not in raw else _dense(raw['betas']).reshape(-1)[:N_SHAPE],
        'shapedirs': _dense(raw['shapedirs'])[:, :, :N_SHAPE],
        'posedirs': posedirs,
        'v_template': _dense(raw['v_template']),
        'J_regressor': _dense(raw['J_regressor']),
        'weights': _dense(raw['weights']),
        'f': _dense(raw['f']).astype(np.uint32),
        'hands_components': _dense(raw['hands_components']),
        'hands_mean': _dense(raw['hands_mean']).reshape(-1),
        'kintree_table': _dense(raw['kintree_table']).astype(np.int64),
    }
    return out


def resolve_mano(mano_dir: str | None, seed: int | None = None) -> tuple[dict, str]:
    """Real constants from ``<mano_dir>/MANO_RIGHT.pkl`` (as the reference, ``manolayer.py:61-65``).  A random synthetic hand is
    used ONLY when the caller asks for one (``seed`` given: tests, benchmarks): silently training on a made-up hand because an
    asset path was wrong would be a quiet failure, so a missing asset raises."""
    if mano_dir:
        for cand in (os.path.join(mano_dir, 'MANO_RIGHT.pkl'), os.path.join(mano_dir, 'models', 'MANO_RIGHT.pkl')):
            if os.path.isfile(cand):
                return load_mano_pkl(cand), cand
    if seed is None:
        raise FileNotFoundError(f'MANO_RIGHT.pkl not found under {mano_dir!r}: pass mano_data=..., or synthetic_seed=<int> for a '
                                'synthetic MANO-shaped hand (tests / benchmarks only)')
    return synthetic_mano(seed), f'synthetic(seed={seed})'
