"""CPU oracle for the DeepImpute hot path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference`` legs may
import this package.  Nothing under ``deepimpute_b200/`` imports it; the product path has no CPU fallback.

PARITY STATUS: **parity unpinned** for the neural-network arithmetic.  The reference delegates it to
TensorFlow/Keras (un-vendored, unpinned: reference ``setup.py:20-28``), which cannot be installed here, and the
reference's own tests hold no golden vector for it (``tests/multinet_test.py:30-33`` only runs to completion).
The oracle restates the published Keras 2.x semantics the reference relies on (cited per function) and is
guarded by hand-derived fp64 cases and finite-difference gradient checks in ``tests/test_oracle.py``.
The host-side partitioning IS pinned: ``tests/test_partition_parity.py`` checks it against the reference's own
``filter_genes`` / ``setTargets`` / ``setPredictors`` / ``get_distance_matrix`` imported from ``/root/reference``
and against golden vectors minted from them (``tests/golden/``, ``scripts/make_golden.py``).
"""
