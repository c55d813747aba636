"""CPU oracle for the DGGM / E-DSAM hot path (TEST INFRASTRUCTURE ONLY).

This package is a from-scratch CPU restatement (numpy + torch-CPU library ops)
of the reference's depth-guidance path.  It exists to check the CUDA product
path; nothing under ``rgb-d-instance-segmentation_b200/`` may import it.  Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs use it, and only as the checker / reported baseline.

Parity pin: the reference ships no golden vectors for this path,
so the oracle is pinned against outputs of the reference itself, imported in the
build container by ``oracle/make_golden.py`` (committed) and stored under
``tests/golden/``.  ``tests/test_oracle_golden.py`` replays them.
"""
from .hotpath import *  # noqa: F401,F403
