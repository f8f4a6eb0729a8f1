"""oracle/ — TEST INFRASTRUCTURE ONLY.

CPU restatements of the reference hot path used as the parity checker.  Nothing in
``edvr_b200/`` may import this package; only ``tests/``, ``__graft_entry__.smoke()``
and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` do.
"""
