"""oracle/ -- CPU restatement of the reference hot path.  TEST INFRASTRUCTURE ONLY.

Nothing in the product package (anoddpm_b200/) imports this.  Allowed users:
tests/, __graft_entry__.smoke(), bench.py's cpu_baseline / --impl reference legs.

Parity status: PINNED -- every function here is checked against golden vectors that
were produced by importing the unmodified reference in the build container
(tests/golden/make_golden.py, outputs committed under tests/golden/*.npz).
"""
