"""
CPU oracle for the detprocess per-event optimal-filter hot path.

THIS PACKAGE IS TEST INFRASTRUCTURE.  Only ``tests/``, ``__graft_entry__.smoke()``
and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.
Nothing under ``detprocess_b200/`` imports it; the product path fails loudly when
the CUDA extension is missing instead of falling back to this code.

Parity status (see DESIGN.md "Oracle"):

* window reductions (baseline / integral / maximum / minimum), window indices,
  YAML parsing, channel algebra: restated from files that ARE in the reference
  tree and, for the reductions, pinned against numpy itself (the library the
  reference calls).
* OF1x1 amp / t0 / chi2 / lowchi2 / resolutions: the arithmetic lives in the
  third-party package QETpy (``qetpy>=1.8.6``, reference ``setup.py:73``) which is
  not vendored in the reference tree and not installable here.  The restatement
  follows QETpy's published OptimumFilter algorithm; **upstream parity is
  unpinned** -- the reference ships no tests or golden vectors for it.  The
  oracle is pinned instead by convention-independent known-answer tests
  (``tests/test_oracle_of.py``) and ``oracle/dump_golden.py`` lets anyone with
  QETpy installed close the loop.
"""
