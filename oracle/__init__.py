"""TEST INFRASTRUCTURE ONLY.

`oracle` is the checker for the B200 hot path, never part of it. Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s CPU-baseline / `--impl reference` legs may import it.

  oracle.ref     ctypes binding of oracle/_ref/libsatref.so = the UNMODIFIED reference translation
                 units compiled from /root/reference (+ oracle/ref_harness.cpp wiring). Parity anchor.
  oracle.port    ctypes binding of oracle/liboracle.so = our own C restatement (oracle.c) of the same
                 algorithms, each function citing the reference file:line it follows; pinned against
                 oracle.ref bit-for-bit in tests/test_oracle_*.py and against tests/golden/*.npz.

Parity status: the reference ships NO tests / golden vectors for this path, so the
pin is (a) the reference's own code compiled here (oracle/_ref) and (b) fixtures generated from it
by tests/golden/make_golden.py.
"""
