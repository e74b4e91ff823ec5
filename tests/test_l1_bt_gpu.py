"""GPU parity: logistic level 1 for binary traits (k-fold IRLS and closed-form LOO) vs the oracle,
on level-0 predictors produced by the GPU itself from a synthetic fileset."""
import numpy as np
import pytest

import helpers
from oracle import prep, step1, step1_bt

pytestmark = pytest.mark.gpu


def _problem(tmp_path, loocv):
    pb = helpers.synthetic_problem(tmp_path, N=800, M=400, P=2, C=3, bsize=100, K=4, seed=3, loocv=loocv)
    pr = pb.prep
    rng = np.random.default_rng(5)
    # binary traits with a genetic signal: threshold the (residualised) quantitative ones
    y_raw = (pr.Y + 0.3 * rng.standard_normal(pr.Y.shape) > 0.4).astype(float) * pr.mask
    return pb, y_raw


@pytest.mark.parametrize("loocv", [False, True])
def test_logistic_level1_matches_oracle(tmp_path, loocv):
    pb, y_raw = _problem(tmp_path, loocv)
    pr = pb.prep
    st = pb.gpu_step1()
    nb = len(pb.blocks)
    for b in range(nb):
        pb.gpu_l0_block(st, b)
    assert st.status() == 0
    B = nb * 5
    h1 = prep.set_ridge_params(5)
    tau = B * (1 - h1) / h1 * 3 / np.pi ** 2
    P = y_raw.shape[1]
    off = np.stack([step1_bt.null_offset(y_raw[:, p], pr.X, pr.mask[:, p]) for p in range(P)], axis=1)
    cs, best = st.l1_fit_bt(y_raw, off, np.tile(tau, (P, 1)))
    chr_of_block = [c for c, _, _ in pb.blocks]
    loco = st.loco(chr_of_block)                        # [P, N, 23]
    chrs = sorted(set(chr_of_block))
    chr_cols = [(c, chr_of_block.index(c) * 5, chr_of_block.count(c) * 5) for c in chrs]
    for p in range(P):
        W = np.hstack([st.fetch_W(b, p) for b in range(nb)])
        if loocv:
            cs_o = step1_bt.level1_logistic_loocv(W, y_raw[:, p], off[:, p], pr.mask[:, p], tau)
        else:
            cs_o, betas = step1_bt.level1_logistic_kfold(W, y_raw[:, p], off[:, p], pr.mask[:, p], tau, pb.fold_sizes)
        np.testing.assert_allclose(cs[:, p, :], cs_o, rtol=1e-7, atol=1e-9)
        best_o, _ = step1_bt.output_table(cs_o, pr.neff[p], B, tau)
        assert best[p] == best_o
        if loocv:
            pred = step1_bt.predictions_binary_loocv(W, y_raw[:, p], off[:, p], pr.mask[:, p], tau[best_o], chr_cols)
        else:
            pred = step1_bt.predictions_binary_kfold(W, betas, best_o, pb.fold_sizes, chr_cols)
        ref = step1.loco_matrix(pred, chr_cols)
        np.testing.assert_allclose(loco[p], ref, rtol=1e-6, atol=1e-8)
    st.close()
