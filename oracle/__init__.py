"""oracle -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU restatements of the reference's chi-square log-likelihood path
(/root/reference/C/ChiSquare.c, mda.py:1540-1578) plus loaders for the
unmodified reference library compiled in place (oracle/_ref/).  Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import anything from here, and only as the
checker or the timed CPU baseline -- never on the product path.  Nothing in
``mimic_mda_b200/`` imports this package.

Parity status: PINNED (see chisq_oracle.c header and tests/test_oracle.py).
"""
