"""Host-side phenotype/covariate preparation used by bench.py (numpy, O(N*P*C)).

Mirrors what the C++ driver does before the first block (reference prep_run,
src/Pheno.cpp:1060-1114: orthonormal covariate basis :1660-1681, residualise + scale the
phenotypes :1799-1834, mean-impute missing QT values :1903-1935).  Not a kernel target.
"""
import numpy as np


def prepare_qt(Y, cov, na=None):
    """Y: N x P raw, cov: N x (C-1) covariates (intercept is prepended), na: N x P bool.

    Returns X (N x C orthonormal), Yres (N x P residualised, unit sd), mask (N x P uint8),
    in_analysis (N uint8), neff (P).  All samples are analysed (QT step 1 keeps and
    mean-imputes missing phenotypes).
    """
    N, P = Y.shape
    Y = Y.astype(np.float64).copy()
    if na is not None:
        for p in range(P):
            ok = ~na[:, p]
            Y[~ok, p] = Y[ok, p].sum() / ok.sum()
    X = np.hstack([np.ones((N, 1)), cov])
    d, v = np.linalg.eigh(X.T @ X)
    nz = int((d > d[-1] * 1e-15).sum())
    Xb = (X @ v[:, -nz:]) / np.sqrt(d[-nz:])[None, :]
    Y = Y - Xb @ (Xb.T @ Y)
    neff = np.full(P, float(N))
    Y = Y / (np.linalg.norm(Y, axis=0) / np.sqrt(neff - nz))[None, :]
    return (np.asfortranarray(Xb), np.asfortranarray(Y), np.ones((N, P), dtype=np.uint8, order="F"),
            np.ones(N, dtype=np.uint8), neff)


def fold_sizes(n, k):
    """Contiguous folds when every sample is analysed (src/Data.cpp:407-428)."""
    t = n // k
    s = [t] * k
    s[-1] = n - t * (k - 1)
    return np.array(s, dtype=np.int64)


def ridge_grid(n):
    v = np.arange(n) / (n - 1.0)
    v[0], v[-1] = 0.01, 0.99
    return v
