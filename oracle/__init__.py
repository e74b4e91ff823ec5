"""CPU oracle for the regenie Step-1 / Step-2 hot path.

TEST INFRASTRUCTURE ONLY.  This package restates, in numpy float64, the
algorithm of the reference (rgcgithub/regenie v4.1.2) for the path named in
BASELINE.json `north_star`.  Each function cites the reference file:line it
follows.  Only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` /
`--impl reference` legs of `bench.py` may import it; the product path
(`regenie_b200`) never does and fails loudly without its CUDA library.

Parity status: the QT Step-1/Step-2 functions have no golden vectors in the
reference's own tests -> "parity unpinned" for QT-only
functions; the functions shared with the binary-trait golden run
(bed/bgen decode, level-0 LOOCV ridge, LOCO assembly, BT score test,
approximate Firth, summary-statistics printing) are pinned against
`tests/golden/example/test_bin_out_firth_Y1.regenie` (see
tests/test_oracle_golden.py).
"""
