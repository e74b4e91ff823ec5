"""Phenotype / covariate preparation (oracle; test infrastructure only).

Restates rgcgithub/regenie v4.1.2:
  pheno_read               src/Pheno.cpp:148-364
  covariate_read           src/Pheno.cpp:573-700
  read_pheno_and_cov       src/Pheno.cpp:50-146
  setMasks                 src/Pheno.cpp:810-841
  pheno_impute_miss        src/Pheno.cpp:1903-1935
  prep_run                 src/Pheno.cpp:1060-1114
  getBasis                 src/Pheno.cpp:1660-1681
  residualize_phenotypes   src/Pheno.cpp:1799-1834
  set_folds                src/Data.cpp:401-431
  set_ridge_params         src/Regenie.cpp:1497-1508
"""
from dataclasses import dataclass

import numpy as np

MISSING = -999.0            # src/Regenie.hpp:215
NUMTOL = 1e-6               # src/Regenie.hpp:220
EIG_REL_TOL = 1e-15         # src/Regenie.hpp:227


def convert_double(s: str) -> float:
    """src/Regenie.cpp:1663-1675."""
    if s == "NA" or s in ("nan", "inf"):
        return MISSING
    return float(s)


def read_table(path, sample_index: dict, n: int, name_filter=None):
    """Read a 'FID IID c1 c2 ...' text table into an [n x k] array ordered by sample_index.

    Returns (names, values, present); rows of samples absent from the genotype file are
    ignored (src/Pheno.cpp:216-217), duplicates are an error (:222-226).
    """
    with open(path) as fh:
        hdr = fh.readline().rstrip("\r\n").split()
        if len(hdr) < 2 or hdr[0] != "FID" or hdr[1] != "IID":
            raise ValueError("header must start with: FID IID.")
        names = hdr[2:]
        keep = [i for i, nm in enumerate(names) if name_filter is None or name_filter(nm)]
        vals = np.zeros((n, len(keep)))
        present = np.zeros(n, dtype=bool)
        for line in fh:
            t = line.split()
            if not t:
                continue
            if len(t) != 2 + len(names):
                raise ValueError("incorrectly formatted file.")
            idx = sample_index.get(t[0] + "_" + t[1])
            if idx is None:
                continue
            if present[idx]:
                raise ValueError("individual appears more than once: FID=%s IID=%s" % (t[0], t[1]))
            present[idx] = True
            vals[idx] = [convert_double(t[2 + i]) for i in keep]
    return [names[i] for i in keep], vals, present


def set_ridge_params(n: int) -> np.ndarray:
    """src/Regenie.cpp:1497-1508: linspace(0,1,n) with the ends replaced by 0.01 / 0.99."""
    v = np.arange(n) / (n - 1.0)
    v[0], v[-1] = 0.01, 0.99
    return v


def get_basis(X: np.ndarray):
    """Orthonormal covariate basis, src/Pheno.cpp:1660-1681."""
    d, v = np.linalg.eigh(X.T @ X)
    nz = int((d > d[-1] * EIG_REL_TOL).sum())
    Xb = (X @ v[:, -nz:]) / np.sqrt(d[-nz:])[None, :]
    return Xb, nz


@dataclass
class Prepared:
    """State after read_pheno_and_cov + prep_run (what `struct phenodt` / `filter` hold)."""
    keys: list                 # FID_IID in genotype-file order (after --remove)
    pheno_names: list
    Y: np.ndarray              # N x P residualised + scaled (QT, and BT step 1)
    Y_raw: np.ndarray          # N x P raw (BT) or None
    mask: np.ndarray           # N x P bool (masked_indivs)
    X: np.ndarray              # N x C orthonormal basis, zero rows outside the analysis
    in_analysis: np.ndarray    # N bool
    neff: np.ndarray           # P
    scale_Y: np.ndarray        # P
    ncov: int
    n_analyzed: int
    bt: bool = False


def prepare(keys, pheno_file, covar_file=None, bt=False, step=1, strict=False,
            pheno_filter=None, rint=False) -> Prepared:
    """read_pheno_and_cov + prep_run for Step 1 (QT or BT) and Step 2 QT.

    For Step 2 the caller applies the LOCO-availability mask and re-runs `finish_prep`.
    """
    n = len(keys)
    sidx = {k: i for i, k in enumerate(keys)}
    names, Y, in_ph = read_table(pheno_file, sidx, n, pheno_filter)
    P = len(names)
    strict = strict or P == 1                                   # src/Pheno.cpp:198
    mask = np.ones((n, P), dtype=bool)
    Y_raw = None
    if bt:                                                       # src/Pheno.cpp:296-315
        Y_raw = Y.copy()
        bad = (Y_raw != 0) & (Y_raw != 1)
        if ((Y_raw != MISSING) & bad).any():
            raise ValueError("a phenotype value is not 0/1/NA")
        mask &= ~bad
    miss = Y == MISSING
    if step == 2 and not bt:                                     # rm_missing_qt (src/Regenie.hpp:302)
        mask &= ~miss
    if strict:                                                   # src/Pheno.cpp:333-337
        anym = miss.any(axis=1)
        mask[anym] = False
        all_miss = anym
    else:
        all_miss = miss.all(axis=1)
    in_ph = in_ph & ~all_miss
    mask &= in_ph[:, None]                                       # src/Pheno.cpp:343

    X = np.ones((n, 1))                                          # intercept, src/Pheno.cpp:79
    in_cov = np.ones(n, dtype=bool)
    if covar_file:
        cnames, Cv, in_cov = read_table(covar_file, sidx, n, lambda nm: nm not in names)
        in_cov = in_cov & ~(Cv == MISSING).any(axis=1)           # src/Pheno.cpp:695-698
        X = np.hstack([X, Cv])
    in_an = in_ph & in_cov                                       # src/Pheno.cpp:101
    return _finish(keys, names, Y, Y_raw, mask, X, in_an, bt, step, strict, rint)


def _finish(keys, names, Y, Y_raw, mask, X, in_an, bt, step, strict, rint=False):
    # setMasks, src/Pheno.cpp:810-841
    in_an = in_an & (mask.all(axis=1) if strict else mask.any(axis=1))
    mask = mask & in_an[:, None]
    Y = Y * in_an[:, None]
    if Y_raw is not None:
        Y_raw = Y_raw * in_an[:, None]
    X = X * in_an[:, None]
    n_analyzed = int(in_an.sum())
    neff = mask.sum(axis=0).astype(float)
    if rint and not bt:                                          # apply_rint / rint_pheno, src/Pheno.cpp:1937-2010
        from scipy.stats import norm, rankdata
        for j in range(Y.shape[1]):
            sel = (Y[:, j] != MISSING) & mask[:, j]
            r = rankdata(Y[sel, j], method="average")
            Y[sel, j] = norm.ppf((r - 3 / 8.0) / (sel.sum() - 2 * 3 / 8.0 + 1))

    # pheno_impute_miss, src/Pheno.cpp:1903-1935
    if (not bt) or step == 1:
        for j in range(Y.shape[1]):
            y = Y[:, j]
            if not bt:
                ok = y != MISSING
                tot = y[ok].sum()
                ns = (in_an & ok).sum()
                y[~ok] = tot / ns
            else:
                m = mask[:, j]
                y[~m] = y[m].sum() / m.sum()
        Y = Y * mask

    # prep_run: orthonormal basis then residualise + scale, src/Pheno.cpp:1104-1175
    Xb, ncov = get_basis(X)
    scale_Y = np.ones(Y.shape[1])
    if (not bt) or step == 1:
        beta = Y.T @ Xb                                          # P x C
        Y = Y - (Xb @ beta.T) * mask
        scale_Y = np.linalg.norm(Y, axis=0) / np.sqrt(neff - ncov)
        if scale_Y.min() < NUMTOL:
            raise ValueError("phenotype has sd=0.")
        Y = Y / scale_Y[None, :]
    return Prepared(list(keys), names, Y, Y_raw, mask, Xb, in_an, neff, scale_Y, ncov,
                    n_analyzed, bt)


def set_folds(in_analysis: np.ndarray, k: int) -> np.ndarray:
    """Contiguous fold sizes over genotype-file order (src/Data.cpp:401-431)."""
    n = len(in_analysis)
    target = int(np.floor(in_analysis.sum() / k))
    if target < 1:
        raise ValueError("not enough samples are present for %d-fold CV." % k)
    sizes = np.ones(k, dtype=np.int64)
    n_non_miss, cum, cur = 0, 0, 0
    for i in range(n):
        if in_analysis[i]:
            n_non_miss += 1
        if n_non_miss == target:
            sizes[cur] = i - cum + 1
            cum += sizes[cur]
            n_non_miss, cur = 0, cur + 1
        elif cur == k - 1:
            sizes[cur] = n - i
            break
    return sizes


def set_blocks(chrom: np.ndarray, bsize: int):
    """Blocks never straddle chromosomes (src/Data.cpp:311-334, get_block_size :579-586).

    Returns a list of (chrom, start, size) over the SNP index range.
    """
    blocks = []
    chrs = []
    for c in chrom:
        if not chrs or chrs[-1] != c:
            chrs.append(int(c))
    start = 0
    for c in chrs:
        n_c = int((chrom == c).sum())
        nb = int(np.ceil(n_c / bsize))
        for b in range(nb):
            bs = n_c - b * bsize if (b + 1) * bsize > n_c else bsize
            blocks.append((c, start + b * bsize, bs))
        start += n_c
    return blocks
