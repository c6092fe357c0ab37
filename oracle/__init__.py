"""CPU oracle for the STLPose HRNet keypoint hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in ``stlpose_b200/`` may import this package; only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference``
legs use it, and there only as the checker or the timed CPU baseline -- never as the product.

Parity status: the reference (angelvillar96/STLPose) ships no tests and no golden vectors
(SURVEY.md section 4), so the oracle is pinned against outputs of the reference's own code,
imported in the build container through ``oracle/ref_shim.py`` and frozen as fixtures under
``tests/golden/`` by ``oracle/make_golden.py``.
"""
