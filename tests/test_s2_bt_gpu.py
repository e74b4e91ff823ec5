"""GPU parity: Step-2 binary-trait score test + approximate Firth on BGEN 8-bit dosages vs the oracle.

The data are the reference's own example fileset (tests/golden/example, committed fixtures); the oracle on
exactly this configuration reproduces the reference's golden output file (tests/test_oracle_golden.py).
"""
import math

import numpy as np
import pytest

from oracle import bgen, prep, step2_bt

pytestmark = pytest.mark.gpu

TOL = 1e-5   # north_star: <= 1e-5 relative on BETA / SE / CHISQ


def _load(golden_dir, pheno_cols=1):
    d = golden_dir
    rm = {"_".join(l.split()[:2]) for l in open(d + "/fid_iid_to_remove.txt") if l.strip()}
    b = bgen.Bgen(d + "/example.bgen")
    keep = np.array([k not in rm for k in b.sample_ids])
    keys = [k for k in b.sample_ids if k not in rm]
    pr = prep.prepare(keys, d + "/phenotype_bin.txt", d + "/covariates.txt", bt=True, step=2)
    probs, miss = [], []
    for chrom, pos, rsid, alleles, p0, p1, m in b.variants():
        probs.append(np.stack([p0, p1], axis=1))
        miss.append(np.where(m, 0x82, 0x02).astype(np.uint8))
    return b, keep, keys, pr, np.stack(probs), np.stack(miss)


def _run(pr, keep, probs, miss, traits, blup, z_thr, n_var, block=400):
    from regenie_b200 import capi
    N = len(keep[keep])
    P = len(traits)
    mask = pr.mask[:, traits]
    Y = pr.Y_raw[:, traits]
    sts = [step2_bt.BtChrom(Y[:, j], pr.X, blup[:, j], mask[:, j]) for j in range(P)]
    s2 = capi.Step2(pr.X, mask, pr.in_analysis, pr.n_analyzed, block)
    s2.set_chr_bt(np.stack([s.gamma_sqrt_mask for s in sts], 1), np.stack([s.gamma_sqrt for s in sts], 1),
                  np.stack([s.yres for s in sts], 1), [s.Xg for s in sts], Y,
                  np.stack([s.cov_blup_offset for s in sts], 1))
    sample_idx = np.nonzero(keep)[0].astype(np.int32)
    n_firth = n_fast = n_rows = 0
    worst = 0.0
    for v0 in range(0, n_var, block):
        v1 = min(n_var, v0 + block)
        o = s2.block_bgen8_bt(probs[v0:v1], miss[v0:v1], sample_idx=sample_idx, min_mac=5.0)
        sel = [(i, j) for i in range(v1 - v0) for j in range(P)
               if not (o["flags"][i] & 1) and o["mac"][i, j] >= 5.0 and abs(o["stat"][i, j]) > z_thr]
        fb, fse, flrt, fst = s2.firth([a for a, _ in sel], [b for _, b in sel])
        fmap = {k: n for n, k in enumerate(sel)}
        for i in range(v1 - v0):
            g, iv = bgen.dosage(probs[v0 + i][keep, 0], probs[v0 + i][keep, 1], (miss[v0 + i][keep] & 0x80) != 0)
            for j in range(P):
                r = step2_bt.score_bt(g, iv, pr.in_analysis, mask[:, j], Y[:, j], sts[j], z_thr, N)
                ignored = bool(o["flags"][i] & 1) or o["mac"][i, j] < 5.0 or bool(o["flags"][i] & 16)
                assert (r is None) == ignored, (v0 + i, j)
                if r is None:
                    continue
                n_rows += 1
                assert o["ns"][i, j] == r["n"]
                assert bool(o["flags"][i] & 8) == r["flipped"]
                for k, key in (("af", "af"), ("info", "info"), ("stat", "stat")):
                    assert abs(o[k][i, j] - r[key]) <= TOL * max(abs(r[key]), 1e-3), (v0 + i, j, k, o[k][i, j], r[key])
                if abs(r["stat"]) <= z_thr:
                    got = (o["beta"][i, j], o["se"][i, j], o["chisq"][i, j])
                else:
                    n = fmap[(i, j)]
                    n_firth += 1
                    n_fast += bool(fst[n] & 256)
                    assert (fst[n] & 15) == int(r["test_fail"]), (v0 + i, j, fst[n])
                    if r["test_fail"]:
                        continue
                    got = (fb[n], fse[n], flrt[n])
                for a, c in zip(got, (r["beta"], r["se"], r["chisq"])):
                    rel = abs(a - c) / max(abs(c), 1e-8)
                    worst = max(worst, rel)
                    assert rel <= TOL, (v0 + i, j, got, r)
    s2.close()
    return n_rows, n_firth, n_fast, worst


def test_bt_score_and_firth_golden_configuration(golden_dir):
    """The configuration of the reference's golden file: 1 trait, LOCO = 0, --pThresh 0.01."""
    b, keep, keys, pr, probs, miss = _load(golden_dir)
    z_thr = math.sqrt(6.634896601021213)
    n_rows, n_firth, n_fast, worst = _run(pr, keep, probs, miss, [0], np.zeros((len(keys), 1)), z_thr, 1000)
    assert n_rows == 1000 and n_firth == 20


def test_bt_firth_many_variants_two_traits_missing_dosages(golden_dir):
    """Firth on every variant with |z| > 0.5, two traits with different masks, a non-zero LOCO offset,
    missing dosages, and a flipped block (covers the carriers-only shortcut and both solvers)."""
    b, keep, keys, pr, probs, miss = _load(golden_dir)
    rng = np.random.default_rng(11)
    probs = probs[:600].copy(); miss = miss[:600].copy()
    miss[rng.random(miss.shape) < 0.01] |= 0x80
    # make some variants common-allele-coded so that flip_geno triggers
    probs[::7, :, 0] = np.where(probs[::7, :, 0] + probs[::7, :, 1] <= 255, 255 - probs[::7, :, 0] - probs[::7, :, 1], 0)
    # rare variants (MAC < 50): the reference then iterates over the carriers only
    rare = rng.random(probs[1::5, :, 0].shape) < 0.93
    probs[1::5][rare] = 0
    blup = 0.3 * rng.standard_normal((len(keys), 2))
    n_rows, n_firth, n_fast, worst = _run(pr, keep, probs, miss, [0, 1], blup, 0.5, 600, block=256)
    assert n_firth > 200 and n_fast > 5


def test_bt_spa_matches_oracle(golden_dir):
    """Saddlepoint approximation (rg_s2_spa) on every variant with |z| > 0.5: dense and sparse ("fast") variants, two
    traits, missing dosages, flipped alleles; chisq / LOG10P / BETA vs the oracle restatement of run_SPA_test_snp."""
    from regenie_b200 import capi
    b, keep, keys, pr, probs, miss = _load(golden_dir)
    rng = np.random.default_rng(12)
    probs = probs[:500].copy(); miss = miss[:500].copy()
    miss[rng.random(miss.shape) < 0.01] |= 0x80
    probs[::7, :, 0] = np.where(probs[::7, :, 0] + probs[::7, :, 1] <= 255, 255 - probs[::7, :, 0] - probs[::7, :, 1], 0)
    rare = rng.random(probs[1::3, :, 0].shape) < 0.8
    probs[1::3][rare] = 0                                              # sparse genotypes -> fast SPA
    blup = 0.3 * rng.standard_normal((len(keys), 2))
    N, P = len(keys), 2
    mask, Y = pr.mask[:, :2], pr.Y_raw[:, :2]
    sts = [step2_bt.BtChrom(Y[:, j], pr.X, blup[:, j], mask[:, j]) for j in range(P)]
    s2 = capi.Step2(pr.X, mask, pr.in_analysis, pr.n_analyzed, 250)
    s2.set_chr_bt(np.stack([s.gamma_sqrt_mask for s in sts], 1), np.stack([s.gamma_sqrt for s in sts], 1),
                  np.stack([s.yres for s in sts], 1), [s.Xg for s in sts], Y, None,
                  np.stack([s.phat for s in sts], 1))
    sample_idx = np.nonzero(keep)[0].astype(np.int32)
    z_thr = 0.5
    n_spa = n_fast = n_fail = 0
    for v0 in range(0, 500, 250):
        o = s2.block_bgen8_bt(probs[v0:v0 + 250], miss[v0:v0 + 250], sample_idx=sample_idx, min_mac=5.0)
        sel = [(i, j) for i in range(250) for j in range(P)
               if not (o["flags"][i] & 17) and o["mac"][i, j] >= 5.0 and abs(o["stat"][i, j]) > z_thr]
        pv, status = s2.spa([a for a, _ in sel], [c for _, c in sel])
        for n, (i, j) in enumerate(sel):
            g, iv = bgen.dosage(probs[v0 + i][keep, 0], probs[v0 + i][keep, 1], (miss[v0 + i][keep] & 0x80) != 0)
            r = step2_bt.score_bt(g, iv, pr.in_analysis, mask[:, j], Y[:, j], sts[j], z_thr, N, correction="spa")
            assert r is not None and abs(r["stat"]) > z_thr
            n_spa += 1
            n_fast += bool(status[n] & 256)
            assert bool(status[n] & 15) == r["test_fail"], (v0 + i, j, status[n])
            if r["test_fail"]:
                n_fail += 1
                continue
            pval = max(step2_bt.NL_DBL_DMIN, pv[n])
            chisq = step2_bt.chisq1_from_pvalue(pval)
            assert abs(chisq - r["chisq"]) <= TOL * r["chisq"], (v0 + i, j, chisq, r["chisq"])
            assert abs(-math.log10(pval) - r["logp"]) <= TOL * max(r["logp"], 1e-3)
            beta = math.copysign(1.0, o["beta"][i, j]) * math.sqrt(chisq) * o["se"][i, j]
            assert abs(beta - r["beta"]) <= TOL * abs(r["beta"])
    s2.close()
    assert n_spa > 300 and n_fast > 30


def test_bt_on_2bit_rows_matches_oracle_and_dosage_path(tmp_path):
    """rg_s2_block_bed_bt (tensor-core sums on 2-bit rows) vs the oracle and vs the 8-bit dosage path fed the same hard
    calls: statistics, flip / sparse flags, Firth and SPA on the resident block; N and A1FREQ bit for bit."""
    from regenie_b200 import capi, synth
    import helpers
    from oracle import plink
    N, M, P = 900, 256, 2
    g = synth.genotypes(N, M, seed=31, miss=0.02, maf_hi=0.5)
    g[::5] = np.where(g[::5] == 3, 3, 2 - np.minimum(g[::5], 2))           # some variants coded on the major allele (flip)
    g[1::4][(np.random.default_rng(1).random(g[1::4].shape) < 0.9) & (g[1::4] != 3)] = 0   # rare / sparse
    Yq, cov, na = synth.phenotypes(np.where(g == 3, 0, g).astype(np.uint8), P, 3, seed=31, na_frac=0.03)
    prefix = helpers.write_fileset(str(tmp_path), g, (Yq > 0.3).astype(float), cov, na)
    keys, _ = plink.read_fam(prefix + ".fam")
    bim = plink.read_bim(prefix + ".bim")
    pr = prep.prepare(keys, str(tmp_path) + "/pheno.txt", str(tmp_path) + "/covar.txt", bt=True, step=2)
    mask, Y = pr.mask, pr.Y_raw
    rng = np.random.default_rng(3)
    blup = 0.2 * rng.standard_normal((N, P))
    sts = [step2_bt.BtChrom(Y[:, j], pr.X, blup[:, j], mask[:, j]) for j in range(P)]
    s2 = capi.Step2(pr.X, mask, pr.in_analysis, pr.n_analyzed, 256)
    s2.set_chr_bt(np.stack([s.gamma_sqrt_mask for s in sts], 1), np.stack([s.gamma_sqrt for s in sts], 1),
                  np.stack([s.yres for s in sts], 1), [s.Xg for s in sts], Y,
                  np.stack([s.cov_blup_offset for s in sts], 1), np.stack([s.phat for s in sts], 1))
    packed = plink.read_bed_rows(prefix + ".bed", N, bim.offset)
    o = s2.block_bed_bt(packed)
    z_thr = 1.0
    sel = [(i, j) for i in range(M) for j in range(P)
           if not (o["flags"][i] & 17) and o["mac"][i, j] >= 5.0 and abs(o["stat"][i, j]) > z_thr]
    fb, fse, flrt, fst = s2.firth([a for a, _ in sel], [c for _, c in sel])
    pv, sst = s2.spa([a for a, _ in sel], [c for _, c in sel])
    # the dosage path on the same calls
    graw = plink.decode_bed(packed, N)
    probs = np.zeros((M, N, 2), dtype=np.uint8)
    probs[:, :, 0] = (graw == 2) * 255
    probs[:, :, 1] = (graw == 1) * 255
    miss = np.where(graw == -3, 0x82, 0x02).astype(np.uint8)
    od = s2.block_bgen8_bt(probs, miss)
    for k in ("ns", "flags"):
        assert np.array_equal(o[k], od[k]), k
    assert np.array_equal(o["af"], od["af"])                                   # bit for bit
    for k in ("stat", "beta", "se", "mac"):
        np.testing.assert_allclose(o[k], od[k], rtol=1e-9, atol=1e-12)
    fmap = {k: n for n, k in enumerate(sel)}
    n_chk = 0
    for i in range(M):
        for j in range(P):
            r = step2_bt.score_bt(graw[i], np.zeros(N), pr.in_analysis, mask[:, j], Y[:, j], sts[j], z_thr, N)
            ignored = bool(o["flags"][i] & 17) or o["mac"][i, j] < 5.0
            assert (r is None) == ignored, (i, j)
            if r is None:
                continue
            assert o["ns"][i, j] == r["n"] and o["af"][i, j] == r["af"]
            assert abs(o["stat"][i, j] - r["stat"]) <= TOL * max(1.0, abs(r["stat"]))
            if abs(r["stat"]) > z_thr and not r["test_fail"]:
                n = fmap[(i, j)]
                assert (fst[n] & 15) == 0
                for a, c in zip((fb[n], fse[n], flrt[n]), (r["beta"], r["se"], r["chisq"])):
                    assert abs(a - c) <= TOL * max(abs(c), 1e-8), (i, j, a, c)
                n_chk += 1
    assert n_chk > 50 and (sst & 15).max() == 0 and np.isfinite(pv).all()
    s2.close()
