"""Binary-trait Step 1 (oracle only; test infrastructure).

Needed to PIN the oracle against the reference's own known answers: the only golden vectors the
reference ships sit downstream of a `--bt` Step 1 (test/test_bash.sh:58-89 greps `0.4504` on the
`min value` line; example/test_bin_out_firth_Y1.regenie is the Step-2 output of that fit).
Level 0 is shared with the QT path (ridge_level_0_loocv), so a match here pins the bed decode,
the phenotype/covariate prep, the level-0 LOOCV ridge and the selection/printing logic.

Restates rgcgithub/regenie v4.1.2:
  get_pvec / get_wvec / get_logist_dev   src/Step1_Models.cpp:1760-1828
  fit_logistic                           src/Step1_Models.cpp:156-222
  fit_null_logistic (step-1 use)         src/Step1_Models.cpp:54-140  (called from src/Pheno.cpp:1608)
  run_log_ridge_loocv                    src/Step1_Models.cpp:1288-1375
  ridge_logistic_level_1_loocv           src/Step1_Models.cpp:1159-1286
  ridge_logistic_level_1 (k-fold)        src/Step1_Models.cpp:966-1157  (make_predictions_binary src/Data.cpp:1346-1428)
  make_predictions_binary_loocv          src/Data.cpp:1484-1573
  Data::output (non-QT criterion)        src/Data.cpp:1025-1077
"""
import numpy as np

NUMTOL = 1e-6
NUMTOL_EPS = 10 * np.finfo(float).eps      # src/Regenie.hpp:225
NITER_MAX = 50                             # :335
NITER_LS = 25                              # :338
NITER_RIDGE = 100                          # :287
L1_RIDGE_TOL = 1e-4                        # :289
L1_RIDGE_EPS = 1e-5                        # :290
TOL = 1e-8                                 # :226


def get_pvec(eta, eps=NUMTOL_EPS):
    """src/Step1_Models.cpp:1797-1804."""
    p = 1.0 - 1.0 / (np.exp(np.clip(eta, -700, 700)) + 1.0)
    p = np.where(eta > 30.0, 1.0 / (1.0 + eps), p)
    p = np.where(eta < -30.0, eps / (1.0 + eps), p)
    return p


def logist_dev(y, p, mask):
    """-2 log-lik, src/Step1_Models.cpp:1819-1828."""
    ll = np.where(y == 0, np.log(1 - p), np.log(p))
    return -2.0 * ll[mask].sum()


def fit_logistic(y, X, offset, mask, beta, check_hs_dev, numtol=NUMTOL):
    """IRLS with step halving (src/Step1_Models.cpp:156-222).  Returns (ok, beta, eta, p)."""
    eta = offset + X @ beta
    p = get_pvec(eta)
    dev_old = logist_dev(y, p, mask)
    m = mask.astype(float)
    diff_dev = 0.0
    betanew = beta.copy()
    it = 0
    small_score = False
    while it < NITER_MAX:
        it += 1
        w = np.where(mask, p * (1 - p), 1.0)
        if (w == 0).any():
            return False, beta, eta, p
        XtW = X.T * (w * m)
        z = np.where(mask, eta - offset + (y - p) / w, 0.0)
        betanew = np.linalg.solve(XtW @ X, XtW @ z)
        ok = False
        for _ in range(NITER_LS):
            eta = offset + X @ betanew
            p = get_pvec(eta)
            dev_new = logist_dev(y, p, mask)
            if ((p[mask] > 0) & (p[mask] < 1)).all() and ((not check_hs_dev) or dev_new < dev_old):
                ok = True
                break
            betanew = (beta + betanew) / 2
        if not ok:
            return False, beta, eta, p
        score = X.T @ np.where(mask, y - p, 0.0)
        smax = np.abs(score).max()
        if smax < numtol:
            break
        if (not small_score) and it < 20 and smax < 1:
            small_score = True
        if small_score and it > 20 and smax > 5:
            return False, beta, eta, p
        diff_dev = abs(dev_new - dev_old) / (0.1 + abs(dev_new))
        beta = betanew
        dev_old = dev_new
    else:
        it += 1
    if ((diff_dev == 0) or (diff_dev >= numtol)) and it > NITER_MAX:
        return False, beta, eta, p
    return True, betanew, eta, p


def null_offset(y_raw, X, mask):
    """fit_null_logistic in Step 1: covariate-only fit, returns the linear predictor (offset_nullreg)."""
    beta0 = np.zeros(X.shape[1])
    zero = np.zeros(len(y_raw))
    for chk in (True, False):
        ok, b, eta, p = fit_logistic(y_raw, X, zero, mask, beta0.copy(), chk)
        if ok:
            return eta
    raise ValueError("logistic regression did not converge")


def run_log_ridge_loocv(lam, beta, y, X, offset, mask):
    """Penalised logistic Newton with Cholesky (src/Step1_Models.cpp:1288-1375)."""
    m = mask.astype(float)
    B = X.shape[1]
    eta = offset + X @ beta
    p = get_pvec(eta)
    fn_start = logist_dev(y, p, mask) + lam * (beta ** 2).sum()
    w = np.where(mask, p * (1 - p), 1.0)
    score = X.T @ np.where(mask, y - p, 0.0) - lam * beta
    betanew = beta
    dev_conv = False
    it = 0
    converged_by_score = False
    while it < NITER_RIDGE:
        it += 1
        H = lam * np.eye(B) + (X.T * (w * m)) @ X
        step = np.linalg.solve(H, score)
        for _ in range(NITER_LS):
            betanew = beta + step
            eta = offset + X @ betanew
            p = get_pvec(eta)
            fn_end = logist_dev(y, p, mask) + lam * (betanew ** 2).sum()
            w = np.where(mask, p * (1 - p), 1.0)
            if fn_end < fn_start + NUMTOL:
                break
            step = step / 2
        score = X.T @ np.where(mask, y - p, 0.0) - lam * betanew
        dev_conv = abs(fn_end - fn_start) / (0.01 + abs(fn_end)) < TOL
        if np.abs(score).max() < L1_RIDGE_TOL:
            converged_by_score = True
            break
        beta = betanew
        fn_start = fn_end
    if (not converged_by_score) and (not dev_conv):
        return False, betanew, p, w
    return True, betanew, p, w


def level1_logistic_loocv(W, y_raw, offset, mask, tau):
    """ridge_logistic_level_1_loocv for one phenotype -> cumsum [6 x R1] (Sx,Sy,Sx2,Sy2,Sxy,-LL)."""
    m = mask.astype(float)
    B = W.shape[1]
    cs = np.zeros((6, len(tau)))
    beta = np.zeros(B)
    for j, t in enumerate(tau):
        ok, beta, p, w = run_log_ridge_loocv(t, beta, y_raw, W, offset, mask)   # warm starts, :1203-1213
        if not ok:
            raise ValueError("ridge logistic regression did not converge")
        H = t * np.eye(B) + (W.T * (w * m)) @ W
        V1 = np.linalg.solve(H, W.T)                                            # B x N
        v2 = (W * V1.T).sum(axis=1) * w
        b_loo = beta[:, None] - V1 * ((y_raw - p) / (1 - v2))[None, :]         # :1250-1253
        pred = (W * b_loo.T).sum(axis=1) + offset
        p1 = 1 - 1 / (np.exp(pred) + 1)
        p1 = np.clip(p1, L1_RIDGE_EPS, 1 - L1_RIDGE_EPS)
        sel = mask
        yy = y_raw[sel]; pp = p1[sel]
        cs[0, j] = pp.sum(); cs[1, j] = yy.sum(); cs[2, j] = (pp ** 2).sum(); cs[3, j] = (yy ** 2).sum()
        cs[4, j] = (pp * yy).sum()
        cs[5, j] = -np.where(yy == 0, np.log(1 - pp), np.log(pp)).sum()
    return cs


def predictions_binary_loocv(W, y_raw, offset, mask, tau_best, chr_cols):
    """make_predictions_binary_loocv (src/Data.cpp:1484-1573): refit at tau*, LOO betas, per-chr dot."""
    m = mask.astype(float)
    B = W.shape[1]
    ok, beta, p, w = run_log_ridge_loocv(tau_best, np.zeros(B), y_raw, W, offset, mask)
    H = tau_best * np.eye(B) + (W.T * (w * m)) @ W
    V1 = np.linalg.solve(H, W.T)
    v2 = (W * V1.T).sum(axis=1) * w
    bfin = beta[:, None] - V1 * ((y_raw - p) / (1 - v2))[None, :]
    pred = np.zeros((W.shape[0], len(chr_cols)))
    for ci, (_, ctr, nn) in enumerate(chr_cols):
        pred[:, ci] = (W[:, ctr:ctr + nn] * bfin[ctr:ctr + nn].T).sum(axis=1)
    return pred


def output_table(cs, neff, B, tau):
    """Rows of the Rsq/MSE/-logLik table (src/Data.cpp:1054-1074) and the argmin of -logLik/N."""
    perf = cs[5] / neff
    best = 0
    mv = 1e10
    for j, v in enumerate(perf):
        if v < mv:
            best, mv = j, v
    rows = []
    for j in range(len(tau)):
        h = B / (B + (np.pi ** 2 / 3) * tau[j])
        num = cs[4, j] - cs[0, j] * cs[1, j] / neff
        rsq = num * num / ((cs[2, j] - cs[0, j] ** 2 / neff) * (cs[3, j] - cs[1, j] ** 2 / neff))
        sse = cs[2, j] + cs[3, j] - 2 * cs[4, j]
        rows.append("  %5s : Rsq = %g, MSE = %g, -logLik/N = %g%s" % ("%g" % h, rsq, sse / neff, cs[5, j] / neff,
                                                                   "<- min value" if j == best else ""))
    return best, rows


def level1_logistic_kfold(W, y_raw, offset, mask, tau, fold_sizes):
    """ridge_logistic_level_1, k-fold branch (src/Step1_Models.cpp:966-1157, !within_sample_l0).

    Per fold i and ridge value j (warm starts over j): IRLS on the samples outside fold i until
    max|score| < l1_ridge_tol, then the CV sums over the masked samples of fold i.
    Returns (cumsum [6 x R1], beta [K][B x R1])."""
    N, B = W.shape
    starts = np.concatenate([[0], np.cumsum(fold_sizes)])
    K = len(fold_sizes)
    cs = np.zeros((6, len(tau)))
    betas = []
    for i in range(K):
        test = np.zeros(N, dtype=bool)
        test[starts[i]:starts[i + 1]] = True
        train = mask & ~test
        m = train.astype(float)
        bnew = np.zeros(B)
        bi = np.zeros((B, len(tau)))
        for j, t in enumerate(tau):
            bold = bnew
            it = 0
            converged = False
            while it < NITER_RIDGE:
                it += 1
                eta = offset + W @ bold
                p = get_pvec(eta)
                w = np.where(train, p * (1 - p), 1.0)
                if (w == 0).any():
                    raise ValueError("Zeros occurred in Var(Y) during ridge logistic regression")
                z = np.where(train, (eta - offset) + (y_raw - p) / w, 0.0)
                XtW = W.T * (w * m)
                bnew = np.linalg.solve(t * np.eye(B) + XtW @ W, XtW @ z)
                for _ in range(NITER_LS):                         # halve only while some weight is exactly 0
                    p = get_pvec(offset + W @ bnew)
                    if not (np.where(train, p * (1 - p), 1.0) == 0).any():
                        break
                    bnew = (bold + bnew) / 2
                p = get_pvec(offset + W @ bnew)
                score = W.T @ np.where(train, y_raw - p, 0.0) - t * bnew
                if np.abs(score).max() < L1_RIDGE_TOL:
                    converged = True
                    break
                bold = bnew
            if not converged:
                raise ValueError("Penalized logistic regression did not converge")
            bi[:, j] = bnew
            sel = mask & test
            etat = offset[sel] + W[sel] @ bnew
            p1 = np.clip(1 - 1 / (np.exp(etat) + 1), L1_RIDGE_EPS, 1 - L1_RIDGE_EPS)
            yy = y_raw[sel]
            cs[0, j] += p1.sum(); cs[1, j] += yy.sum(); cs[2, j] += (p1 ** 2).sum(); cs[3, j] += (yy ** 2).sum()
            cs[4, j] += (p1 * yy).sum()
            cs[5, j] += -np.where(yy == 0, np.log(1 - p1), np.log(p1)).sum()
        betas.append(bi)
    return cs, betas


def predictions_binary_kfold(W, betas, best, fold_sizes, chr_cols):
    """make_predictions_binary, !within_sample_l0 (src/Data.cpp:1400-1416): out-of-fold per-chromosome dot products."""
    N = W.shape[0]
    starts = np.concatenate([[0], np.cumsum(fold_sizes)])
    pred = np.zeros((N, len(chr_cols)))
    for i in range(len(fold_sizes)):
        rows = slice(starts[i], starts[i + 1])
        for ci, (_, ctr, nn) in enumerate(chr_cols):
            pred[rows, ci] = W[rows, ctr:ctr + nn] @ betas[i][ctr:ctr + nn, best]
    return pred
