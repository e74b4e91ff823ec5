"""Pin the oracle against the reference's OWN known answers (no GPU).

(1) test/test_bash.sh:58-89 runs `--step 1 --bed example/example --exclude snplist_rm.txt --covarFile
    covariates.txt --phenoFile phenotype_bin.txt --remove fid_iid_to_remove.txt --bsize 100 --bt --lowmem`
    and requires a log line containing both `0.4504` and `min value`.  N_analyzed = 494 < 5000, so the
    reference silently runs LOOCV (src/Data.cpp:353-356).  Reproducing that line exercises, end to end,
    the .bed decode + --exclude/--remove handling, phenotype/covariate prep, the level-0 LOOCV ridge
    shared with the QT path, the logistic level-1 LOOCV and the selection/printing rules.
"""
import numpy as np

from oracle import plink, prep, step1, step1_bt


def bt_step1_tables(d, with_loco=False):
    excl = {l.split()[0] for l in open(d + "/snplist_rm.txt") if l.strip()}
    rm = {"_".join(l.split()[:2]) for l in open(d + "/fid_iid_to_remove.txt") if l.strip()}
    bim = plink.read_bim(d + "/example.bim", exclude=excl)
    keys_file, _ = plink.read_fam(d + "/example.fam")
    keep = np.array([k not in rm for k in keys_file])
    keys = [k for k in keys_file if k not in rm]
    pr = prep.prepare(keys, d + "/phenotype_bin.txt", d + "/covariates.txt", bt=True, step=1)
    assert pr.n_analyzed == 494 and len(bim.ids) == 994
    blocks = prep.set_blocks(bim.chrom, 100)
    packed = plink.read_bed_rows(d + "/example.bed", len(keys_file), bim.offset)
    h = prep.set_ridge_params(5)
    lam = len(bim.ids) * (1 - h) / h
    cols = [[] for _ in range(2)]
    for c, s, bs in blocks:
        g = plink.decode_bed(packed[s:s + bs], len(keys_file), keep=keep)
        gi, _ = plink.mean_impute_block(g, pr.in_analysis)
        Gt, _ = step1.residualize_genotypes(gi, pr.X, pr.in_analysis, pr.n_analyzed, pr.ncov)
        W = step1.level0_loocv(Gt, pr.Y, pr.mask, lam, pr.neff)
        for ph in range(2):
            cols[ph].append(W[ph])
    out = []
    for ph in range(2):
        W = np.hstack(cols[ph])
        B = W.shape[1]
        tau = B * (1 - h) / h * 3 / np.pi ** 2                      # src/Step1_Models.cpp:2115-2117
        off = step1_bt.null_offset(pr.Y_raw[:, ph], pr.X, pr.mask[:, ph])
        cs = step1_bt.level1_logistic_loocv(W, pr.Y_raw[:, ph], off, pr.mask[:, ph], tau)
        best, rows = step1_bt.output_table(cs, pr.neff[ph], B, tau)
        if with_loco:
            chr_cols = [(1, 0, B)]                                   # example/ holds chromosome 1 only
            pred = step1_bt.predictions_binary_loocv(W, pr.Y_raw[:, ph], off, pr.mask[:, ph], tau[best], chr_cols)
            out.append((best, rows, step1.loco_matrix(pred, chr_cols), keys, pr))
        else:
            out.append((best, rows))
    return out


def test_reference_known_answer_0_4504(golden_dir):
    tables = bt_step1_tables(golden_dir)
    lines = [r for _, rows in tables for r in rows if "min value" in r]
    assert len(lines) == 2
    assert any("0.4504" in l for l in lines), lines          # test/test_bash.sh:87


def test_reference_golden_step2_bt_firth_file(golden_dir):
    """(2) example/test_bin_out_firth_Y1.regenie -- the one golden output file the reference ships
    (docs/docs/options.md:20-51: Step 2 on example.bgen, --bt --firth --approx --pThresh 0.01, --remove,
    covariates, LOCO from the BT Step 1 above).  example/ has a single chromosome, so the LOCO row for
    chr 1 is identically 0 (src/Data.cpp:1847-1858) and the file pins Step 2 on its own: BGEN v1.2 decode,
    A1FREQ / INFO / N, minor-allele flip, sparse/dense switch, BT score test, approximate Firth (20 of the
    1000 rows), LOG10P and the native row format.  Text columns must match exactly; numeric columns to
    the printed 6 significant digits (the reference is built with -ffast-math)."""
    import math

    from oracle import bgen, step2, step2_bt
    d = golden_dir
    rm = {"_".join(l.split()[:2]) for l in open(d + "/fid_iid_to_remove.txt") if l.strip()}
    b = bgen.Bgen(d + "/example.bgen")
    keep = np.array([k not in rm for k in b.sample_ids])
    keys = [k for k in b.sample_ids if k not in rm]
    pr = prep.prepare(keys, d + "/phenotype_bin.txt", d + "/covariates.txt", bt=True, step=2)
    y, mask = pr.Y_raw[:, 0], pr.mask[:, 0]
    st = step2_bt.BtChrom(y, pr.X, np.zeros(len(keys)), mask)
    z_thr = math.sqrt(6.634896601021213)                  # chi2_1 quantile(1 - pThresh = 0.99), src/Data.cpp:2116-2120
    gold = [l.split() for l in open(d + "/test_bin_out_firth_Y1.regenie")]
    assert " ".join(gold[0]) == step2.HEADER_INFO.strip()
    gold = gold[1:]
    gi = n_firth = 0
    for chrom, pos, rsid, alleles, p0, p1, miss in b.variants():
        g, iv = bgen.dosage(p0[keep], p1[keep], miss[keep])
        r = step2_bt.score_bt(g, iv, pr.in_analysis, mask, y, st, z_thr, len(keys))
        if r is None:
            continue
        row = step2.sumstats_row(int(chrom), pos, rsid, alleles[1], alleles[0], r["af"], r["n"], r["beta"], r["se"],
                                 r["chisq"], r["logp"], info=r["info"], test_pass=not r["test_fail"]).split()
        ref = gold[gi]
        gi += 1
        n_firth += abs(r["stat"]) > z_thr
        assert row[:9] == ref[:9], (row, ref)             # CHROM GENPOS ID ALLELE0 ALLELE1 A1FREQ INFO N TEST
        for a, c in zip(row[9:13], ref[9:13]):            # BETA SE CHISQ LOG10P
            assert abs(float(a) - float(c)) <= 1e-4 * abs(float(c)), (row, ref)
        assert row[13] == ref[13]
    assert gi == len(gold) == 1000 and n_firth == 20
