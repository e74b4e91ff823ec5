"""The quantitative-trait arithmetic has no golden output in the reference: the k-fold / LOOCV ridge levels and
compute_score_qt are pinned only by restating the code.  This file checks those restatements against the MODEL the code
implements, computed by different algorithms from the textbook definitions (Mbatchou et al. 2021, Methods):

* level 0 / level 1 ridge, k-fold: out-of-fold predictions of  argmin |y - G'b|^2 + lambda |b|^2  through an augmented
  least-squares problem (np.linalg.lstsq on [G' ; sqrt(lambda) I]) instead of the eigendecomposition of G G';
* LOOCV: every sample actually left out and the ridge refitted, instead of the hat-matrix shortcut;
* Step-2 score test: chi-square = (n - C) * partial r^2 and BETA = the OLS coefficient of the genotype in  y ~ covariates + g,
  from np.linalg.lstsq on the raw text inputs (Frisch-Waugh), instead of projections on the orthonormal basis.

It is test infrastructure on top of test infrastructure: nothing here is on the product path.  Tolerances are those of a
well-conditioned double-precision solve (1e-8)."""
import numpy as np
import pytest

import helpers
from oracle import plink, prep, step1, step2
from regenie_b200 import synth


def ridge_lstsq(A, y, lam):
    """argmin |y - A b|^2 + lam |b|^2 as one least-squares problem (A: n x p)."""
    p = A.shape[1]
    Aa = np.vstack([A, np.sqrt(lam) * np.eye(p)])
    ya = np.concatenate([y, np.zeros(p)])
    return np.linalg.lstsq(Aa, ya, rcond=None)[0]


def rel(a, b):
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-300))


@pytest.fixture(scope="module")
def pb(tmp_path_factory):
    d = tmp_path_factory.mktemp("fp")
    return helpers.synthetic_problem(d, N=240, M=64, P=2, C=3, bsize=32, miss=0.0, K=4, na_frac=0.0, drop=False)


def test_level0_kfold_is_out_of_fold_ridge(pb):
    """ridge_level_0 (src/Step1_Models.cpp:458-613): for every fold the predictor of its samples comes from a ridge fit on the
    other folds; the columns are then centred and scaled to unit variance (:539-557)."""
    W_o, _, _, G = pb.oracle_l0(0)                       # G: residualised, scaled genotypes of block 0 (bs x N)
    pr = pb.prep
    starts = np.concatenate([[0], np.cumsum(pb.fold_sizes)])
    for ph in range(pr.Y.shape[1]):
        y = pr.Y[:, ph]
        cols = []
        for lam in pb.lam:
            pred = np.zeros(len(y))
            for f in range(len(pb.fold_sizes)):
                s, e = starts[f], starts[f + 1]
                out = np.r_[0:s, e:len(y)]
                b = ridge_lstsq(G[:, out].T, y[out], lam)
                pred[s:e] = G[:, s:e].T @ b
            pred = pred * pr.mask[:, ph]
            n = pr.neff[ph]
            mean = pred.sum() / n
            sd = np.sqrt(((pred ** 2).sum() - n * mean ** 2) / (n - 1))
            cols.append((pred - mean) / sd)
        assert rel(np.stack(cols, axis=1), W_o[ph]) < 1e-8


def test_level1_kfold_sums_are_out_of_fold_ridge(pb):
    """ridge_level_1 (src/Step1_Models.cpp:790-873): the five running sums behind the MSE / Rsq table."""
    W_o, _, _, _ = pb.oracle_l0(0)
    W2, _, _, _ = pb.oracle_l0(1)
    pr = pb.prep
    W = np.hstack([W_o[0], W2[0]])                      # N x B stacked predictors of phenotype 0
    y = pr.Y[:, 0]
    tau = np.array([0.5, 3.0, 20.0]) * W.shape[1]
    cs, betas = step1.level1_kfold(W, y, pb.fold_sizes, tau)
    starts = np.concatenate([[0], np.cumsum(pb.fold_sizes)])
    want = np.zeros_like(cs)
    for f in range(len(pb.fold_sizes)):
        s, e = starts[f], starts[f + 1]
        out = np.r_[0:s, e:len(y)]
        for j, t in enumerate(tau):
            b = ridge_lstsq(W[out], y[out], t)
            assert rel(b, betas[f][:, j]) < 1e-8
            p1 = W[s:e] @ b
            want[0, j] += p1.sum(); want[2, j] += p1 @ p1; want[4, j] += p1 @ y[s:e]
        want[1] += y[s:e].sum(); want[3] += y[s:e] @ y[s:e]
    assert np.allclose(cs, want, rtol=1e-8, atol=1e-9)


def test_loocv_shortcuts_equal_leaving_every_sample_out(pb):
    """ridge_level_0_loocv / ridge_level_1_loocv (src/Step1_Models.cpp:615-726, :875-963) use (pred - h y) / (1 - h) with the
    hat-matrix diagonal h; here every sample is really left out and the ridge refitted."""
    _, _, _, G = pb.oracle_l0(0)
    pr = pb.prep
    n = G.shape[1]
    y = pr.Y[:, 0]
    lam = pb.lam[2]
    loo = np.array([G[:, i] @ ridge_lstsq(np.delete(G, i, axis=1).T, np.delete(y, i), lam) for i in range(n)])
    W = step1.level0_loocv(G, pr.Y, pr.mask, np.array([lam]), pr.neff)[0][:, 0]
    m = loo * pr.mask[:, 0]
    nf = pr.neff[0]
    mean = m.sum() / nf
    assert rel((m - mean) / np.sqrt(((m ** 2).sum() - nf * mean ** 2) / (nf - 1)), W) < 1e-8
    # level 1: the sums over the left-out predictions
    Wl = np.hstack([w for w in pb.oracle_l0(0)[0]])[:, :6]
    tau = np.array([2.0, 30.0])
    cs = step1.level1_loocv(Wl, y, tau, pr.neff[0], pr.ncov)
    for j, t in enumerate(tau):
        p = np.array([Wl[i] @ ridge_lstsq(np.delete(Wl, i, axis=0), np.delete(y, i), t) for i in range(n)])
        assert np.allclose([cs[0, j], cs[2, j], cs[4, j]], [p.sum(), p @ p, p @ y], rtol=1e-8, atol=1e-9)


def test_score_test_is_partial_correlation_and_ols_beta(tmp_path):
    """compute_score_qt (src/Step2_Models.cpp:343-440) with --ignore-pred, one trait, no missing values: CHISQ = (n - C) r^2 with
    r the partial correlation of phenotype and genotype given the covariates, BETA = the OLS coefficient of the genotype -
    both from least squares on the numbers in the text files."""
    g = synth.genotypes(300, 40, seed=5, miss=0.0)
    Y, cov, na = synth.phenotypes(g, 1, 3, seed=5)
    prefix = helpers.write_fileset(str(tmp_path), g, Y, cov, na)
    keys, _ = plink.read_fam(prefix + ".fam")
    pr = prep.prepare(keys, str(tmp_path) + "/pheno.txt", str(tmp_path) + "/covar.txt", step=2)
    n, C = pr.n_analyzed, pr.ncov
    res, p_sd, scf = step2.compute_res(pr.Y, np.zeros_like(pr.Y), pr.mask, pr.neff, pr.ncov, pr.scale_Y)
    YtX = res.T @ pr.X
    bim = plink.read_bim(prefix + ".bim")
    rows = plink.read_bed_rows(prefix + ".bed", len(keys), bim.offset)
    y_txt = np.array([float(l.split()[2]) for l in open(str(tmp_path) + "/pheno.txt").read().splitlines()[1:]])
    c_txt = np.array([[float(x) for x in l.split()[2:]] for l in open(str(tmp_path) + "/covar.txt").read().splitlines()[1:]])
    A0 = np.hstack([np.ones((len(y_txt), 1)), c_txt])
    rss0 = np.sum((y_txt - A0 @ np.linalg.lstsq(A0, y_txt, rcond=None)[0]) ** 2)
    checked = {True: 0, False: 0}                       # both branches of check_sparse_G (src/Geno.cpp:3180-3212) are met
    for i in range(len(bim.ids)):
        graw = plink.decode_bed(rows[i:i + 1], len(keys))[0]
        vs = step2.variant_stats(graw, pr.in_analysis, pr.mask)
        if vs["ignored"]:
            continue
        gg = vs["g"]
        sc = step2.score_qt(gg, pr.X, res, pr.mask, pr.in_analysis, n, C, scf, YtX, True)
        A1 = np.hstack([A0, gg[:, None]])
        b1 = np.linalg.lstsq(A1, y_txt, rcond=None)[0]
        r2 = 1.0 - np.sum((y_txt - A1 @ b1) ** 2) / rss0
        assert abs(sc["chisq"][0] - (n - C) * r2) <= 1e-8 * max(1.0, (n - C) * r2)
        assert abs(sc["beta"][0] - b1[-1]) <= 1e-8 * max(1.0, abs(b1[-1])), (i, sc["beta"], b1[-1])
        checked[bool(sc["is_sparse"])] += 1
    assert checked[True] >= 5 and checked[False] >= 5, checked
