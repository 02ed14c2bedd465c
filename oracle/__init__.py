"""oracle/ -- CPU restatement of tapir's Townsend-PI hot path.  TEST INFRASTRUCTURE ONLY.

Nothing under ``tapir_b200/`` imports this package.  Only ``tests/``, ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs and ``__graft_entry__.smoke()`` may use it, and
only as the checker (or as the timed CPU baseline), never as the product path.

Parity status: PINNED.
  * ``townsend_port`` (numpy + scipy.integrate.quad, the reference's own third-party
    arithmetic) is checked against golden vectors produced by importing the reference's
    unmodified ``tapir/compute.py`` (tests/golden/make_golden.py) and against the reference's
    own fixtures (test-R-townsend-output.npy, test-30-50-integral.npy, the PhyDesign
    3/4-decimal pins of tapir/tests/test_compute.py:54-71).
  * ``quadpack_port.c`` (plain C QUADPACK dqagse/dqk21/dqelg/dqpsrt restatement) is checked
    bit-for-bit (result, abserr, neval) against scipy.integrate.quad driven with the same
    integrand, over single- and multi-panel cases (tests/test_oracle_quadpack.py).
"""
