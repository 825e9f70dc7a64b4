"""CPU oracle for the optimizer-step hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import this package.  The
product path (``multioptpy_b200``) never does and fails loudly when its CUDA
library is missing.

Parity status: PINNED.  ``oracle/np_oracle.py`` is a NumPy restatement of the
reference algorithm; it is checked (``tests/test_oracle_golden.py``) against
golden vectors in ``tests/golden/`` that were produced by running the
unmodified reference itself (``oracle/gen_golden.py`` via ``oracle/ref_shim.py``)
in the build container.  The reference ships no tests or golden vectors of
its own.
"""
