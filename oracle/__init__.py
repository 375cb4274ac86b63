"""CPU oracle for the MHEntropy multi-hypothesis hot path.  TEST INFRASTRUCTURE ONLY.

A plain-PyTorch (CPU, fp32 or fp64) restatement of the reference algorithm for the path
SURVEY.md §8 scopes: the conditional RealNVP flow (``hand/flows.py``), the MANO layer
(``hand/manopth/manolayer.py`` + ``hand/ManoLayer.py``) and the reprojection / Laplace / prior /
entropy reductions of ``hand/network.py``.  Every function cites the reference ``file:line`` it
follows.

Pinning: the reference holds no golden vectors or known-answer tests for this path,
so the oracle is pinned against outputs of the reference itself, executed unmodified in the build
container through ``oracle/ref_shim.py``; the resulting fixtures live in ``tests/golden/`` together
with the script that made them (``tests/golden/make_golden.py``).  ``tests/test_oracle_vs_golden.py``
checks the oracle against every fixture and, where ``/root/reference`` is present, against the
live reference.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import this package — always as the checker or as the timed CPU baseline,
never as part of the product path.  ``mhentropy_b200`` never imports it.
"""
