"""The compiled Eigen/OpenMP restatement (oracle/ref_eigen, built against the reference's vendored Eigen 3.4.0) and
the numpy oracle are two independent restatements of the same reference functions; they must agree.  This gives the
QT k-fold level-0 arithmetic (which has no golden vector in the reference's tests) a second pin that uses
the reference's own SelfAdjointEigenSolver, and the Step-2 QT score test a second implementation."""
import numpy as np
import pytest

import helpers
from oracle import plink, ref_eigen, step2


@pytest.mark.parametrize("seed,N,M,bsize", [(3, 1200, 160, 80), (11, 2051, 130, 130)])
def test_level0_kfold_eigen_matches_numpy_oracle(tmp_path, seed, N, M, bsize):
    pb = helpers.synthetic_problem(tmp_path, N=N, M=M, P=3, C=3, bsize=bsize, miss=0.02, seed=seed)
    pr = pb.prep
    for b in range(len(pb.blocks)):
        c, s, bs = pb.blocks[b]
        W_np, _, _, _ = pb.oracle_l0(b)
        W_e, phases = ref_eigen.l0_block_kfold(pb.packed[s:s + bs], pb.n_file, pr.in_analysis, pr.X, pr.Y, pr.mask,
                                               pb.fold_sizes, pb.lam, pr.neff, pr.n_analyzed, threads=2)
        assert (phases >= 0).all()
        for ph in range(3):
            err = np.abs(W_e[ph] - W_np[ph]).max() / np.abs(W_np[ph]).max()
            assert err < 1e-9, (b, ph, err)


def test_level0_eigen_is_thread_count_invariant_to_rounding(tmp_path):
    pb = helpers.synthetic_problem(tmp_path, N=900, M=64, P=2, C=3, bsize=64, seed=5)
    pr = pb.prep
    c, s, bs = pb.blocks[0]
    args = (pb.packed[s:s + bs], pb.n_file, pr.in_analysis, pr.X, pr.Y, pr.mask, pb.fold_sizes, pb.lam, pr.neff, pr.n_analyzed)
    W1, _ = ref_eigen.l0_block_kfold(*args, threads=1)
    W4, _ = ref_eigen.l0_block_kfold(*args, threads=4)
    for a, b in zip(W1, W4):
        assert np.abs(a - b).max() < 1e-10


def test_step2_qt_eigen_matches_numpy_oracle(tmp_path):
    pb = helpers.synthetic_problem(tmp_path, N=1500, M=120, P=3, C=3, bsize=120, miss=0.03, seed=9)
    pr = pb.prep
    rng = np.random.default_rng(1)
    res = rng.normal(size=(pb.n_file, 3)) * pr.mask
    res /= np.linalg.norm(res, axis=0) / np.sqrt(pr.neff - pr.ncov)
    scf = np.array([1.3, 0.7, 2.0])
    YtX = res.T @ pr.X
    out = ref_eigen.s2_block_qt_bed(pb.packed, pb.n_file, pr.in_analysis, pr.X, res, pr.mask, YtX, scf, pr.n_analyzed, threads=2)
    n_checked = 0
    for i in range(pb.M):
        graw = plink.decode_bed(pb.packed[i:i + 1], pb.n_file)[0]
        vs = step2.variant_stats(graw, pr.in_analysis, pr.mask)
        if vs["ignored"]:
            assert out[i, 0] == -1
            continue
        sc = step2.score_qt(vs["g"], pr.X, res, pr.mask, pr.in_analysis, pr.n_analyzed, pr.ncov, scf, YtX, False)
        assert out[i, 0] == vs["af1"] and out[i, 1] == vs["ns1"] and out[i, 3] == sc["is_sparse"]
        for ph in range(3):
            q = out[i, 4 + 5 * ph: 9 + 5 * ph]
            assert q[0] == vs["af"][ph] and q[1] == vs["ns"][ph]
            np.testing.assert_allclose(q[2:], [sc["beta"][ph], sc["se"][ph], sc["chisq"][ph]], rtol=1e-9)
        n_checked += 1
    assert n_checked > 50
