"""Step-1 stacked ridge regression (oracle; test infrastructure only).

numpy-float64 restatement of rgcgithub/regenie v4.1.2 (Eigen 3.4.0 arithmetic):
  Data::residualize_genotypes   src/Data.cpp:190-228
  Data::calc_cv_matrices        src/Data.cpp:729-776
  ridge_level_0                 src/Step1_Models.cpp:458-613
  ridge_level_0_loocv           src/Step1_Models.cpp:615-726
  check_l0 (tau grid)           src/Step1_Models.cpp:2105-2118
  ridge_level_1                 src/Step1_Models.cpp:772-872
  ridge_level_1_loocv           src/Step1_Models.cpp:875-963
  Data::output (tau* choice)    src/Data.cpp:1023-1084
  Data::make_predictions        src/Data.cpp:1196-1267
  Data::make_predictions_loocv  src/Data.cpp:1269-1343
  Data::write_predictions       src/Data.cpp:1795-1982 (LOCO assembly + text)

`SelfAdjointEigenSolver` is restated with numpy.linalg.eigh (LAPACK dsyevd): same
mathematical object, ~1e-13 relative agreement.
"""
import numpy as np

from .plink import NCHROM
from .prep import NUMTOL


def residualize_genotypes(G, X, in_analysis, n_analyzed, ncov):
    """src/Data.cpp:190-228.  G: bs x N (already mean-imputed).  Returns (G~, scale_G)."""
    G = G * in_analysis[None, :]
    beta = G @ X
    G = G - beta @ X.T
    scale = np.linalg.norm(G, axis=1) / np.sqrt(n_analyzed - ncov)
    if scale.min() < NUMTOL:
        raise ValueError("SNP %d has low variance (=%g)." % (int(scale.argmin()), scale.min()))
    return G / scale[:, None], scale


def level0_kfold(G, Y, mask, fold_sizes, lam, neff):
    """calc_cv_matrices (k-fold branch) + ridge_level_0.

    G: bs x N residualised/scaled; Y: N x P; mask: N x P bool; lam: R ridge values
    (already lambda = M(1-h)/h, src/Data.cpp:607).
    Returns W: [P][N x R] level-0 predictors of this block (centred/scaled, :539-557).
    """
    bs, N = G.shape
    P = Y.shape[1]
    R = len(lam)
    starts = np.concatenate([[0], np.cumsum(fold_sizes)])
    GtY_f, GG_f = [], []
    GGt = np.zeros((bs, bs)); GTY = np.zeros((bs, P))
    for f in range(len(fold_sizes)):                       # src/Data.cpp:741-751
        s, e = starts[f], starts[f + 1]
        Gf = G[:, s:e]
        GtY_f.append(Gf @ Y[s:e]); GTY += GtY_f[-1]
        GG_f.append(Gf @ Gf.T); GGt += GG_f[-1]
    pred_all = np.zeros((R, P, N))
    p_sum = np.zeros((R, P)); p_sum2 = np.zeros((R, P))
    for f in range(len(fold_sizes)):                       # src/Step1_Models.cpp:480-527
        s, e = starts[f], starts[f + 1]
        d, V = np.linalg.eigh(GGt - GG_f[f])
        ww2 = V.T @ (GTY - GtY_f[f])
        for j in range(R):
            beta = V @ (ww2 / (d + lam[j])[:, None])       # :494
            pred = (beta.T @ G[:, s:e]) * mask[s:e].T       # :503  (P x Nf)
            p_sum[j] += pred.sum(axis=1)
            p_sum2[j] += (pred ** 2).sum(axis=1)
            pred_all[j, :, s:e] = pred
    W = []
    for ph in range(P):                                    # :539-557
        mean = p_sum[:, ph] / neff[ph]
        invsd = np.sqrt((neff[ph] - 1) / (p_sum2[:, ph] - neff[ph] * mean ** 2))
        W.append(((pred_all[:, ph, :].T - mean[None, :]) * invsd[None, :]))
    return W


def level0_loocv(G, Y, mask, lam, neff):
    """calc_cv_matrices (LOOCV branch, src/Data.cpp:753-768) + ridge_level_0_loocv."""
    bs, N = G.shape
    P = Y.shape[1]
    GGt = G @ G.T
    GTY = G @ Y
    d, V = np.linalg.eigh(GGt)
    Wmat = V.T @ GTY                                       # bs x P
    DL_inv = 1.0 / (d[:, None] + lam[None, :])             # bs x R   (:640)
    VtG = V.T @ G                                          # bs x N   (:654)
    h = np.einsum("kr,kn->nr", DL_inv, VtG ** 2)           # N x R    gvec (:659)
    # pred[n, r, p] = (z2^T Wmat - h * y_n) / (1 - h)       (:660-663)
    zw = np.einsum("kn,kr,kp->nrp", VtG, DL_inv, Wmat, optimize=True)
    pred = (zw - h[:, :, None] * Y[:, None, :]) / (1.0 - h)[:, :, None]
    W = []
    for ph in range(P):                                    # :694-706
        w = pred[:, :, ph] * mask[:, ph][:, None]
        mean = w.sum(axis=0) / neff[ph]
        w = (w - mean[None, :]) * mask[:, ph][:, None]
        sd = np.linalg.norm(w, axis=0) / np.sqrt(neff[ph] - 1)
        W.append(w / sd[None, :])
    return W


def level1_kfold(W, y, fold_sizes, tau):
    """ridge_level_1 for one phenotype.  W: N x B, y: N, tau: R1 values (already B(1-h)/h).

    Returns (cumsum[5 x R1] = Sx, Sy, Sx2, Sy2, Sxy; betas: list over folds of B x R1).
    """
    starts = np.concatenate([[0], np.cumsum(fold_sizes)])
    K = len(fold_sizes)
    XtX_f = [W[starts[f]:starts[f + 1]].T @ W[starts[f]:starts[f + 1]] for f in range(K)]
    XtY_f = [W[starts[f]:starts[f + 1]].T @ y[starts[f]:starts[f + 1]] for f in range(K)]
    XtX = sum(XtX_f); XtY = sum(XtY_f)
    cs = np.zeros((5, len(tau)))
    betas = []
    for f in range(K):
        s, e = starts[f], starts[f + 1]
        d, V = np.linalg.eigh(XtX - XtX_f[f])                     # :828
        VtX2 = V.T @ (XtY - XtY_f[f])
        beta = V @ (VtX2[:, None] / (d[:, None] + tau[None, :]))   # :833-835
        betas.append(beta)
        p1 = W[s:e] @ beta                                        # :847
        yy = y[s:e]
        cs[0] += p1.sum(axis=0)
        cs[1] += yy.sum()
        cs[2] += (p1 ** 2).sum(axis=0)
        cs[3] += (yy ** 2).sum()
        cs[4] += (p1 * yy[:, None]).sum(axis=0)
    return cs, betas


def level1_loocv(W, y, tau, neff, ncov):
    """ridge_level_1_loocv for one phenotype (src/Step1_Models.cpp:875-963)."""
    cs = np.zeros((5, len(tau)))
    cs[3] += neff - ncov                                          # :891
    d, V = np.linalg.eigh(W.T @ W)
    z = V.T @ (W.T @ y)
    T = W @ V                                                     # :928
    for j, t in enumerate(tau):
        w = 1.0 / (d + t)
        cal = (T ** 2) @ w                                        # :935
        pred = (T @ (w * z) - cal * y) / (1.0 - cal)              # :936-937
        cs[0, j] += pred.sum()
        cs[2, j] += pred @ pred
        cs[4, j] += pred @ y
    return cs


def pick_tau(cs, neff):
    """Data::output, QT criterion (src/Data.cpp:1025-1037): argmin (Sx2+Sy2-2Sxy)/Neff."""
    perf = (cs[2] + cs[3] - 2 * cs[4]) / neff
    best, mv = 0, 1e10
    for j, v in enumerate(perf):
        if v < mv:
            best, mv = j, v
    return best


def rsq_mse(cs, neff):
    """Rsq / MSE table, src/Data.cpp:1058-1070."""
    num = cs[4] - cs[0] * cs[1] / neff
    rsq = num * num / ((cs[2] - cs[0] ** 2 / neff) * (cs[3] - cs[1] ** 2 / neff))
    mse = (cs[2] + cs[3] - 2 * cs[4]) / neff
    return rsq, mse


def chrom_columns(blocks, R):
    """[(chrom, first_col, ncols)] in chromosome order (src/Data.cpp:1238-1244)."""
    out = []
    ctr = 0
    for c in sorted({b[0] for b in blocks}):
        nn = sum(1 for b in blocks if b[0] == c) * R
        out.append((c, ctr, nn))
        ctr += nn
    return out


def predictions_kfold(W, betas, fold_sizes, best, chr_cols):
    """Data::make_predictions (src/Data.cpp:1238-1254): N x nchr per-chromosome predictions."""
    starts = np.concatenate([[0], np.cumsum(fold_sizes)])
    pred = np.zeros((W.shape[0], len(chr_cols)))
    for ci, (_, ctr, nn) in enumerate(chr_cols):
        for f in range(len(fold_sizes)):
            s, e = starts[f], starts[f + 1]
            pred[s:e, ci] = W[s:e, ctr:ctr + nn] @ betas[f][ctr:ctr + nn, best]
    return pred


def predictions_loocv(W, y, tau_best, chr_cols):
    """Data::make_predictions_loocv (src/Data.cpp:1288-1328)."""
    B = W.shape[1]
    xtx = W.T @ W + tau_best * np.eye(B)
    d, V = np.linalg.eigh(xtx)
    Hinv = (V / d[None, :]) @ V.T                                 # :1298
    b = Hinv @ (W.T @ y)
    yres = y - W @ b
    HX = Hinv @ W.T                                               # B x N  (:1310)
    cal = (W * HX.T).sum(axis=1)
    b0 = b[:, None] - HX * (yres / (1 - cal))[None, :]            # :1312
    pred = np.zeros((W.shape[0], len(chr_cols)))
    for ci, (_, ctr, nn) in enumerate(chr_cols):
        pred[:, ci] = (W[:, ctr:ctr + nn] * b0[ctr:ctr + nn].T).sum(axis=1)   # :1323
    return pred


def loco_matrix(pred, chr_cols):
    """LOCO assembly (src/Data.cpp:1846-1858): N x 23; absent chromosomes get the full sum."""
    loco = np.repeat(pred.sum(axis=1)[:, None], NCHROM, axis=1)
    for ci, (c, _, nn) in enumerate(chr_cols):
        if nn > 0:
            loco[:, c - 1] -= pred[:, ci]
    return loco


def fmt_g(x: float) -> str:
    """`ostream << double` at default precision 6 == printf('%g')."""
    return "%g" % x


def write_loco(path, keys, in_analysis, mask_ph, loco):
    """write_ID_header / write_chr_row (src/Data.cpp:1926-1982): std::map key order."""
    order = sorted(range(len(keys)), key=lambda i: keys[i])
    order = [i for i in order if in_analysis[i]]
    with open(path, "w") as fh:
        fh.write("FID_IID " + "".join(keys[i] + " " for i in order) + "\n")
        for c in range(NCHROM):
            fh.write("%d " % (c + 1) + "".join(
                (fmt_g(loco[i, c]) if mask_ph[i] else "NA") + " " for i in order) + "\n")


def run_step1_qt(G_blocks, blocks, prep, fold_sizes, n_variants, loocv=False,
                 h0=None, h1=None):
    """Driver for the QT Step-1 numerical path, given decoded+imputed genotype blocks.

    G_blocks: iterable of bs x N mean-imputed blocks (one per entry of `blocks`).
    Returns dict(W, cs, best, pred, loco, tau) with per-phenotype lists.
    """
    from .prep import set_ridge_params
    h0 = set_ridge_params(5) if h0 is None else np.asarray(h0, float)
    h1 = set_ridge_params(5) if h1 is None else np.asarray(h1, float)
    lam = n_variants * (1 - h0) / h0                              # src/Data.cpp:607
    P = prep.Y.shape[1]
    Wcols = [[] for _ in range(P)]
    for Gb in G_blocks:
        Gt, _ = residualize_genotypes(Gb, prep.X, prep.in_analysis, prep.n_analyzed, prep.ncov)
        if loocv:
            Wb = level0_loocv(Gt, prep.Y, prep.mask, lam, prep.neff)
        else:
            Wb = level0_kfold(Gt, prep.Y, prep.mask, fold_sizes, lam, prep.neff)
        for ph in range(P):
            Wcols[ph].append(Wb[ph])
    out = dict(W=[], cs=[], best=[], pred=[], loco=[], tau=[])
    chr_cols = chrom_columns(blocks, len(h0))
    for ph in range(P):
        W = np.hstack(Wcols[ph])
        B = W.shape[1]
        tau = B * (1 - h1) / h1                                   # src/Step1_Models.cpp:2115
        y = prep.Y[:, ph]
        if loocv:
            cs = level1_loocv(W, y, tau, prep.neff[ph], prep.ncov)
            best = pick_tau(cs, prep.neff[ph])
            pred = predictions_loocv(W, y, tau[best], chr_cols)
        else:
            cs, betas = level1_kfold(W, y, fold_sizes, tau)
            best = pick_tau(cs, prep.neff[ph])
            pred = predictions_kfold(W, betas, fold_sizes, best, chr_cols)
        out["W"].append(W); out["cs"].append(cs); out["best"].append(best)
        out["pred"].append(pred); out["loco"].append(loco_matrix(pred, chr_cols)); out["tau"].append(tau)
    return out
