"""Step-1 level 0 on 8-bit BGEN dosages: rg_l0_block_dosage_u8 / rg_l0_block_f64 against the oracle
restatement of readChunkFromBGENFileToG_fast (src/Geno.cpp:1574-1699) + residualize_genotypes + ridge_level_0, on
synthetic probability pairs (missing calls, --remove, --ref-first, LOOCV) and on the reference's example.bgen."""
import numpy as np
import pytest

import helpers
from oracle import bgen as obgen
from oracle import plink, prep, step1
from regenie_b200 import capi

pytestmark = pytest.mark.gpu


def _probs(rng, bs, n, miss_frac=0.02):
    p0 = rng.integers(0, 256, size=(bs, n))
    p1 = (rng.random((bs, n)) * (256 - p0)).astype(np.int64)
    certain = rng.random((bs, n)) < 0.7                      # most calls are (nearly) hard calls
    g = rng.integers(0, 3, size=(bs, n))
    p0 = np.where(certain, (g == 2) * 255, p0)
    p1 = np.where(certain, (g == 1) * 255, p1)
    probs = np.stack([p0, p1], axis=2).astype(np.uint8)
    miss = np.where(rng.random((bs, n)) < miss_frac, 0x82, 0x02).astype(np.uint8)
    return probs, miss


def _oracle_W(probs, miss, keep, pr, fold_sizes, lam, loocv, ref_first=False):
    g = np.stack([obgen.dosage(probs[v, :, 0].astype(np.float64), probs[v, :, 1].astype(np.float64), (miss[v] & 0x80) != 0,
                               ref_first=ref_first)[0] for v in range(probs.shape[0])])
    g = g[:, keep]
    gi, _ = plink.mean_impute_block(g, pr.in_analysis)
    Gt, _ = step1.residualize_genotypes(gi, pr.X, pr.in_analysis, pr.n_analyzed, pr.ncov)
    if loocv:
        return step1.level0_loocv(Gt, pr.Y, pr.mask, lam, pr.neff)
    return step1.level0_kfold(Gt, pr.Y, pr.mask, fold_sizes, lam, pr.neff)


@pytest.mark.parametrize("loocv,ref_first,remove", [(False, False, False), (False, True, True), (True, False, True)])
def test_dosage_block_matches_oracle(tmp_path, loocv, ref_first, remove):
    N, bs, P = 1100, 96, 3
    pb = helpers.synthetic_problem(tmp_path, N=N, M=bs, P=P, C=3, bsize=bs, seed=4, loocv=loocv)
    if remove:
        pb = helpers.Problem(str(tmp_path) + "/syn", str(tmp_path) + "/pheno.txt", str(tmp_path) + "/covar.txt", bs, loocv=loocv,
                             remove={"F3_I3", "F700_I700", "F%d_I%d" % (N - 1, N - 1)})
    rng = np.random.default_rng(11)
    probs, miss = _probs(rng, bs, N)
    st = pb.gpu_step1()
    idx = None if pb.keep.all() else pb.sample_idx
    st.l0_block_dosage_u8(probs, miss, 0, sample_idx=idx, ref_first=ref_first)
    assert st.status() == 0
    W_o = _oracle_W(probs, miss, pb.keep, pb.prep, pb.fold_sizes, pb.lam, loocv, ref_first)
    for ph in range(P):
        W = st.fetch_W(0, ph)
        assert np.abs(W - W_o[ph]).max() / np.abs(W_o[ph]).max() < 1e-9
    if not ref_first:
        # the FP64 entry point on the same dosages (what a .pgen dosage track decodes to) gives the same predictors
        g = np.stack([obgen.dosage(probs[v, :, 0].astype(np.float64), probs[v, :, 1].astype(np.float64),
                                   (miss[v] & 0x80) != 0)[0] for v in range(bs)])
        st2 = pb.gpu_step1()
        st2.l0_block_f64(g, 0, sample_idx=idx)
        assert st2.status() == 0
        for ph in range(P):
            np.testing.assert_allclose(st2.fetch_W(0, ph), st.fetch_W(0, ph), rtol=0, atol=1e-10)
        st2.close()
    st.close()


def test_example_bgen_block_matches_oracle(golden_dir):
    """The reference's own example.bgen (500 samples x 1000 variants, zlib, 8 bit): the first 100-variant block."""
    bg = obgen.Bgen(golden_dir + "/example.bgen")
    vs = []
    for i, v in enumerate(bg.variants()):
        if i == 100:
            break
        vs.append(v)
    probs = np.stack([np.stack([v[4], v[5]], axis=1) for v in vs]).astype(np.uint8)
    miss = np.stack([np.where(v[6], 0x82, 0x02) for v in vs]).astype(np.uint8)
    pb = helpers.Problem(golden_dir + "/example", golden_dir + "/phenotype.txt", golden_dir + "/covariates.txt", 100)
    st = pb.gpu_step1()
    st.l0_block_dosage_u8(probs, miss, 0)
    assert st.status() == 0
    W_o = _oracle_W(probs, miss, pb.keep, pb.prep, pb.fold_sizes, pb.lam, False)
    for ph in range(2):
        W = st.fetch_W(0, ph)
        assert np.abs(W - W_o[ph]).max() / np.abs(W_o[ph]).max() < 1e-9
    st.close()
