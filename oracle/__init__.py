"""TEST INFRASTRUCTURE ONLY.  CPU oracle for the PSLD sampling hot path.

Importable from ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs only.  Nothing in ``psld_b200`` imports it.
"""
