"""CPU oracle for the DynMM gated hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is product code: only
``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl
reference`` legs of ``bench.py`` may import it, and only as the checker or the
timed CPU baseline -- never as a fallback for the CUDA path.
"""
