"""CPU oracle for the LS-SPA hot path.  TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is part of the shipped product.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs
of ``bench.py`` may import it, and there only as the checker or as the timed
CPU baseline -- never as a fallback for the CUDA path.

Parity status: PINNED.  ``oracle/make_golden.py`` imports the unmodified
reference from ``/root/reference`` (numpy 2.3.5 / scipy 1.18.1, the versions in
this image) and stores its outputs under ``tests/golden/``; the
``-m "not gpu"`` tests check this restatement against those fixtures and against
the golden numbers quoted in SURVEY.md section 8c.
"""
