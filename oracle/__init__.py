"""ORACLE package — test infrastructure only.

CPU restatements (PyTorch CPU ops / numpy) of the reference algorithms on the hot path, each function citing the
reference file:line it follows.  Pinned against vectors produced by the unmodified reference
(tests/golden/make_golden.py -> tests/golden/*.npz, checked by tests/test_oracle_golden.py).

Allowed importers: tests/, __graft_entry__.smoke(), and bench.py's cpu_baseline / `--impl reference` legs.
The product package (3d-playground_b200/) never imports from here.
"""
