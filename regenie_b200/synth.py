"""Deterministic synthetic PLINK panels: per-SNP MAF ~ U(0.01, 0.5),
calls ~ Binomial(2, MAF), a fraction of missing calls, packed as real .bed rows."""
import numpy as np

SEED = 20260924
# PLINK 2-bit codes (ref-last): dosage 2 -> 00, missing -> 01, 1 -> 10, 0 -> 11
_CODE = np.array([3, 2, 0, 1], dtype=np.uint8)   # index: 0,1,2 dosage, 3 = missing


def pack_bed(g):
    """g: int array [M, N] with values 0/1/2 and 3 for missing -> uint8 [M, ceil(N/4)]."""
    M, N = g.shape
    pad = (-N) % 4
    c = _CODE[g]
    if pad:
        c = np.concatenate([c, np.zeros((M, pad), dtype=np.uint8)], axis=1)
    c = c.reshape(M, -1, 4)
    return (c[:, :, 0] | (c[:, :, 1] << 2) | (c[:, :, 2] << 4) | (c[:, :, 3] << 6)).astype(np.uint8)


def genotypes(N, M, seed=SEED, miss=0.01, maf_lo=0.01, maf_hi=0.5):
    rng = np.random.default_rng(seed)
    maf = rng.uniform(maf_lo, maf_hi, size=M)
    g = rng.binomial(2, maf[:, None], size=(M, N)).astype(np.uint8)
    # no monomorphic SNPs (reference throws on sd < 1e-6, src/Data.cpp:207)
    for i in np.where(g.min(axis=1) == g.max(axis=1))[0]:
        g[i, rng.integers(0, N, size=3)] = [0, 1, 2]
    if miss > 0:
        g[rng.random(size=(M, N)) < miss] = 3
    return g


def phenotypes(g, P, C, seed=SEED, h2=0.2, n_causal=100, na_frac=0.0):
    """Y = G beta + noise, covariates ~ N(0,1) (+ intercept added by the caller)."""
    M, N = g.shape
    rng = np.random.default_rng(seed + 1)
    gg = np.where(g == 3, 0, g).astype(np.float64)
    gg = (gg - gg.mean(axis=1, keepdims=True))
    sd = gg.std(axis=1, keepdims=True); sd[sd == 0] = 1
    gg /= sd
    Y = np.empty((N, P))
    for p in range(P):
        idx = rng.choice(M, size=min(n_causal, M), replace=False)
        b = rng.normal(size=len(idx)) * np.sqrt(h2 / len(idx))
        Y[:, p] = b @ gg[idx] + rng.normal(size=N) * np.sqrt(1 - h2)
    cov = rng.normal(size=(N, C - 1))
    na = rng.random(size=(N, P)) < na_frac
    return Y, cov, na
