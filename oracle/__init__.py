"""CPU oracle for the FISRnet hot path -- TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is product code.  It restates, on the CPU (torch fp32 / fp64,
numpy, cv2), the arithmetic of the reference's hot path so that the CUDA path in
``fisr_b200`` can be checked against it.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it; ``fisr_b200``
never does (tests/test_no_oracle_in_product.py enforces this).

PARITY UNPINNED: the reference (JihyongOh/FISR @ 34d9305) ships no tests, no golden
tensors and no weights, and its arithmetic lives in un-vendored TensorFlow 1.13.1 /
OpenCV 4.2 wheels that cannot be installed here.  The oracle is
therefore a restatement of ``ops.py`` / ``FISRnet.py`` / ``utils.py`` plus the documented
TF-1.13 op semantics, pinned only by the geometry / colour-conversion fixtures the
reference does ship (tests/golden/scene1_*; see tests/test_oracle_fixtures.py).
"""
