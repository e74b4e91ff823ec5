"""GPU parity: non-PAR chrX rule of Step 2 (males coded 0/2 count half towards MAC; MAC = min(mac, 2N - N_males - mac),
reference src/Geno.cpp:2447-2462 and compute_mac :3077-3108) for the .bed QT path and the BGEN BT path vs the oracle."""
import numpy as np
import pytest

import helpers
from oracle import bgen, plink, prep, step2, step2_bt

pytestmark = pytest.mark.gpu


def test_chrx_mac_rule_qt_bed(tmp_path):
    from regenie_b200 import capi, synth
    N, M, P = 600, 256, 2
    g = synth.genotypes(N, M, seed=21, miss=0.03, maf_hi=0.12)
    rng = np.random.default_rng(9)
    male = rng.random(N) < 0.5
    g[:, male] = np.where(g[:, male] == 1, 2, g[:, male])            # males are coded 0/2
    Y, cov, na = synth.phenotypes(g, P, 3, seed=21, na_frac=0.05)
    prefix = helpers.write_fileset(str(tmp_path), g, Y, cov, na)
    bim = plink.read_bim(prefix + ".bim")
    keys, _ = plink.read_fam(prefix + ".fam")
    pr = prep.prepare(keys, str(tmp_path) + "/pheno.txt", str(tmp_path) + "/covar.txt", step=2)
    res, p_sd, scf = step2.compute_res(pr.Y, np.zeros_like(pr.Y), pr.mask, pr.neff, pr.ncov, pr.scale_Y)
    st = capi.Step2(pr.X, pr.mask, pr.in_analysis, pr.n_analyzed, 256)
    st.set_sex(male)
    st.set_chr(res, scf)
    non_par = (np.arange(M) % 3 != 0)
    packed = plink.read_bed_rows(prefix + ".bed", len(keys), bim.offset)
    st.set_non_par(non_par)
    o = st.block_bed(packed, min_mac=40.0)
    o2 = st.block_bed(packed, min_mac=40.0)                          # flags are consumed: autosomal rule again
    graw = plink.decode_bed(packed, len(keys))
    n_diff = 0
    for i in range(M):
        vs = step2.variant_stats(graw[i], pr.in_analysis, pr.mask, male=male, non_par=bool(non_par[i]))
        va = step2.variant_stats(graw[i], pr.in_analysis, pr.mask)
        assert o["mac_all"][i] == vs["mac1"] and np.array_equal(o["mac"][i], vs["mac"])
        assert o2["mac_all"][i] == va["mac1"] and np.array_equal(o2["mac"][i], va["mac"])
        assert np.array_equal(o["ns"][i], vs["ns"])
        n_diff += vs["mac1"] != va["mac1"]
    assert n_diff > 100
    st.close()


def test_chrx_mac_rule_bt_bgen(golden_dir):
    from regenie_b200 import capi
    d = golden_dir
    b = bgen.Bgen(d + "/example.bgen")
    keys = list(b.sample_ids)
    pr = prep.prepare(keys, d + "/phenotype_bin.txt", d + "/covariates.txt", bt=True, step=2)
    rng = np.random.default_rng(4)
    male = rng.random(len(keys)) < 0.45
    probs, miss = [], []
    for chrom, pos, rsid, alleles, p0, p1, m in b.variants():
        probs.append(np.stack([p0, p1], axis=1)); miss.append(np.where(m, 0x82, 0x02).astype(np.uint8))
        if len(probs) == 200:
            break
    probs, miss = np.stack(probs), np.stack(miss)
    mask = pr.mask[:, [0]]
    y = pr.Y_raw[:, [0]]
    stc = step2_bt.BtChrom(y[:, 0], pr.X, np.zeros(len(keys)), mask[:, 0])
    s2 = capi.Step2(pr.X, mask, pr.in_analysis, pr.n_analyzed, 200)
    s2.set_sex(male)
    s2.set_chr_bt(stc.gamma_sqrt_mask[:, None], stc.gamma_sqrt[:, None], stc.yres[:, None], [stc.Xg], y,
                  stc.cov_blup_offset[:, None])
    non_par = np.ones(200, dtype=np.uint8)
    s2.set_non_par(non_par)
    o = s2.block_bgen8_bt(probs, miss, min_mac=150.0)
    n_ign = 0
    for i in range(200):
        g, iv = bgen.dosage(probs[i][:, 0], probs[i][:, 1], (miss[i] & 0x80) != 0)
        step2.MIN_MAC, step2_bt.MIN_MAC = 150.0, 150.0
        try:
            r = step2_bt.score_bt(g, iv, pr.in_analysis, mask[:, 0], y[:, 0], stc, 1e9, len(keys), male=male, non_par=True)
        finally:
            step2.MIN_MAC, step2_bt.MIN_MAC = 5.0, 5.0
        ignored = bool(o["flags"][i] & 1) or o["mac"][i, 0] < 150.0 or bool(o["flags"][i] & 16)
        assert (r is None) == ignored, i
        n_ign += ignored
        if r is not None:
            assert abs(o["stat"][i, 0] - r["stat"]) <= 1e-8 * max(1.0, abs(r["stat"]))
    assert 0 < n_ign < 200
    s2.close()
