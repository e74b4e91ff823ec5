"""CPU oracle for the regenie Step-1 / Step-2 hot path.

TEST INFRASTRUCTURE ONLY.  This package restates, in numpy float64, the
algorithm of the reference (rgcgithub/regenie v4.1.2) for the path named in
BASELINE.json `north_star`.  Each function cites the reference file:line it
follows.  Only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` /
`--impl reference` legs of `bench.py` may import it; the product path
(`regenie_b200`) never does and fails loudly without its CUDA library.

Parity status (tests/test_oracle_golden.py):
  * PINNED on the reference's own known answers:
      - `0.4504 ... <- min value` in the BT Step-1 log (test/test_bash.sh:58-89): .bed decode,
        --exclude/--remove, phenotype/covariate prep, level-0 LOOCV ridge, logistic level-1
        LOOCV, tau selection and table printing;
      - all 1000 rows of example/test_bin_out_firth_Y1.regenie (docs/docs/options.md:20-51):
        BGEN v1.2 decode, A1FREQ / INFO / N, allele flip, sparse/dense switch, BT score test,
        approximate Firth (20 rows), LOG10P, native row format.
      - example/example.pgen decodes to exactly the calls of example/example.bed (the reference ships both
        for the same 1000 x 500 genotypes): .pgen record decoding, .pvar / .psam parsing.
  * "parity unpinned": the QT-only arithmetic (k-fold level 0/1 for QT, compute_score_qt), the k-fold
    logistic level 1 and the saddlepoint approximation (SPA) have no golden vector in the reference's tests
    (SURVEY.md section 8c); they share their readers, prep, level-0 algebra, null models, LOCO assembly and
    printing with the pinned paths.
"""
