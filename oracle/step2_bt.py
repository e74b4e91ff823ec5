"""Step-2 binary-trait score test with approximate Firth fallback (oracle only; test infrastructure).

Restates rgcgithub/regenie v4.1.2 so the oracle can be checked against the ONE golden output file the
reference ships (example/test_bin_out_firth_Y1.regenie, docs/docs/options.md:20-51):
  fit_null_logistic (Step-2 use)     src/Step1_Models.cpp:54-140
  Data::compute_res_bin              src/Data.cpp:2439-2455
  fit_approx_firth_null / fit_firth_nr   src/Step2_Models.cpp:899-983, 1267-1383
  fit_null_firth (cov_blup_offset)   src/Step2_Models.cpp:985-1060
  parseSnpfromBGEN stats + flip      src/Geno.cpp:2186-2413, 3077-3163
  check_sparse_G                     src/Geno.cpp:3165-3178
  compute_score_bt                   src/Step2_Models.cpp:470-556
  check_pval_snp / get_sumstats      src/Step2_Models.cpp:1988-2041
  fit_firth_logistic_snp_fast        src/Step2_Models.cpp:1158-1252
  fit_firth_pseudo / fit_firth (1 SNP)   src/Step2_Models.cpp:1527-1737
"""
import math

import numpy as np

from .prep import get_basis
from .step1_bt import NUMTOL_EPS, L1_RIDGE_EPS, fit_logistic, get_pvec, logist_dev
from .step2 import MIN_MAC, PROP_ZERO_THR, get_logp

NUMTOL = 1e-6
NUMTOL_FIRTH = 2.5e-4       # src/Regenie.hpp:224
MAXSTEP = 5                 # :339
MAXSTEP_NULL = 25           # :340
NITER_FIRTH = 250           # :336
NITER_FIRTH_NULL = 1000     # :337
NITER_LS = 25               # :338
NITER_MAX = 50              # :335


def null_logistic_offset(y, X, offset, mask):
    """fit_null_logistic in test mode: returns (beta, eta, p).  src/Step1_Models.cpp:79-86."""
    b0 = np.zeros(X.shape[1])
    for chk in (True, False):
        ok, b, eta, p = fit_logistic(y, X, offset, mask, b0.copy(), chk)
        if ok:
            return b, eta, p
    raise ValueError("null logistic regression did not converge")


def firth_nr(y, X, offset, mask, beta, maxstep, niter, tol, check_score_inc=True):
    """fit_firth_nr with all columns included, comp_lrt = False (null model).  :1267-1383."""
    m = mask
    score_old = 1e16
    n_inc = 0
    it = 0
    dev_new = 0.0
    while it < niter:
        it += 1
        eta = offset + X @ beta
        p = get_pvec(eta)
        dev_old = logist_dev(y, p, m)
        w = np.where(m, p * (1 - p), 1.0)
        XtW = X.T * np.sqrt(w)
        XtWX = XtW @ XtW.T
        sign, logdet = np.linalg.slogdet(XtWX)
        dev_old -= logdet
        h = (np.linalg.solve(XtWX, XtW) * XtW).sum(axis=0)
        mod_score = X.T @ np.where(m, y - p + h * (0.5 - p), 0.0)
        step = np.linalg.solve(XtWX, mod_score)
        smax = np.abs(mod_score).max()
        if smax < tol and it >= 2:
            break
        n_inc = n_inc + 1 if smax > score_old else 0
        if check_score_inc and n_inc > 25:
            return False, beta
        mx = np.abs(step).max() / maxstep
        if mx > 1:
            step = step / mx
        ok = False
        for ls in range(1, NITER_LS + 1):
            if ls > 1:
                step = step / 2
            bn = beta + step
            p2 = get_pvec(offset + X @ bn)
            dev_new = logist_dev(y, p2, m)
            w2 = np.where(m, p2 * (1 - p2), 1.0)
            XtW2 = X.T * np.sqrt(w2)
            dev_new -= np.linalg.slogdet(XtW2 @ XtW2.T)[1]
            if dev_new < dev_old:
                ok = True
                break
        if not ok:
            return False, beta
        beta = beta + step
        score_old = smax
    else:
        return False, beta
    return True, beta


def firth_pseudo_1snp(dev0, y, g, offset, mask, carriers, beta, niter, tol):
    """fit_firth_pseudo, single SNP (src/Step2_Models.cpp:1527-1641).  Returns (state, beta, se, lrt)."""
    fast = carriers is not None and len(carriers) > 0
    if fast:
        p = get_pvec(offset + g * beta)
        dev_new = logist_dev(y, p, mask)
        dev_nc = dev_new - logist_dev(y[carriers], p[carriers], mask[carriers])
        gm = g[carriers]; yy = y[carriers]; off = offset[carriers]; mk = mask[carriers]
    else:
        gm = np.where(mask, g, 0.0); yy = y; off = offset; mk = mask
        dev_nc = 0.0
    gsq = gm * gm
    it = 0
    b14 = 0.0
    XtWX = 1.0
    dev_new = 0.0
    betanew = beta
    while it < niter:
        it += 1
        p = get_pvec(off + (g[carriers] if fast else g) * beta)
        dev_new = dev_nc + logist_dev(yy, p, mk)
        w = np.where(mk, p * (1 - p), 1.0)
        d = gsq * w
        XtWX = d.sum()
        dev_new -= math.log(XtWX)
        h = d / XtWX
        ystar = yy + h * (0.5 - p)
        score = (gm * (ystar - p)).sum()
        if abs(score) < tol and it >= 2:
            break
        if it == 14:
            b14 = beta
        if it == 15 and abs(beta - b14) > 0.1:
            return 1, beta, 0.0, 0.0
        nl = 0
        bdiff = 1e16
        while nl < 25:
            nl += 1
            step = score / XtWX
            bnew = abs(step)
            if bnew > bdiff:
                return 2, beta, 0.0, 0.0
            mx = bnew / 5.0
            betanew = beta + (step / mx if mx > 1 else step)
            p = get_pvec(off + (g[carriers] if fast else g) * betanew)
            score = (gm * (ystar - p)).sum()
            if abs(score) < tol:
                break
            w = np.where(mk, p * (1 - p), 1.0)
            if (w == 0).any():
                return 3, beta, 0.0, 0.0
            XtWX = (gsq * w).sum()
            beta = betanew
            bdiff = bnew
        else:
            nl += 1
        if nl > NITER_MAX:
            return 1, beta, 0.0, 0.0
        beta = betanew
    else:
        return 1, beta, 0.0, 0.0
    lrt = dev0 - dev_new
    if lrt < 0:
        return 4, beta, 0.0, lrt
    return 0, beta, math.sqrt(1 / XtWX), lrt


def firth_nr_1snp(dev0, y, g, offset, mask, carriers, beta, maxstep, niter, tol):
    """fit_firth, single SNP Newton-Raphson (src/Step2_Models.cpp:1644-1737)."""
    fast = carriers is not None and len(carriers) > 0
    p = get_pvec(offset + g * beta)
    dev_old = logist_dev(y, p, mask)
    if fast:
        gm = g[carriers]; yy = y[carriers]; off = offset[carriers]; mk = mask[carriers]; gg = g[carriers]
        p = get_pvec(off + gg * beta)
        dev_nc = dev_old - logist_dev(yy, p, mk)
    else:
        gm = np.where(mask, g, 0.0); yy = y; off = offset; mk = mask; gg = g
        dev_nc = 0.0
    w = np.where(mk, p * (1 - p), 1.0)
    gsq = gm * gm
    d = gsq * w
    XtWX = d.sum()
    dev_old -= math.log(XtWX)
    it = 0
    dev_new = dev_old
    while it < niter:
        it += 1
        h = d / XtWX
        score = (gm * (yy - p + h * (0.5 - p))).sum()
        if abs(score) < tol and it >= 2:
            break
        step = score / XtWX
        mx = abs(step) / maxstep
        if mx > 1:
            step /= mx
        ok = False
        for ls in range(1, NITER_LS + 1):
            if ls > 1:
                step /= 2
            bn = beta + step
            p = get_pvec(off + gg * bn)
            dev_new = dev_nc + logist_dev(yy, p, mk)
            w = np.where(mk, p * (1 - p), 1.0)
            d = gsq * w
            XtWX = d.sum()
            dev_new -= math.log(XtWX)
            if dev_new < dev_old:
                ok = True
                break
        if not ok:
            step += 1e-6
        beta += step
        dev_old = dev_new
    else:
        return False, beta, 0.0, 0.0
    lrt = dev0 - dev_new
    if lrt < 0:
        return False, beta, 0.0, lrt
    return True, beta, math.sqrt(1 / XtWX), lrt


class BtChrom:
    """Per-chromosome null state of one binary trait (compute_res_bin + fit_null_firth)."""

    def __init__(self, y_raw, X, blup, mask):
        loco = blup * mask
        self.beta0, eta, p = null_logistic_offset(y_raw, X, loco, mask)
        w = np.where(mask, p * (1 - p), 1.0)                       # get_wvec, src/Step1_Models.cpp:1760
        self.gamma_sqrt = np.sqrt(w)
        self.gamma_sqrt_mask = self.gamma_sqrt * mask
        self.Xg, _ = get_basis(self.gamma_sqrt_mask[:, None] * X)  # X_Gamma, :130-131
        self.yres = (y_raw - p) / self.gamma_sqrt * mask           # src/Data.cpp:2443-2445
        self.phat = p                                              # m_ests.Y_hat_p, src/Step1_Models.cpp:128
        # null approximate Firth: covariate effects become an offset (src/Step2_Models.cpp:899-983, 1014-1017)
        ok, bf = firth_nr(y_raw, X, blup, mask, self.beta0.copy(), MAXSTEP_NULL, NITER_FIRTH_NULL, 50 * NUMTOL)
        if not ok:
            raise ValueError("null Firth did not converge")
        self.cov_blup_offset = X @ bf + blup


# ------------------------------------------------------------------------------------------------ saddlepoint (SPA)
TOL_SPA = float(np.finfo(float).eps) ** 0.25        # src/Regenie.hpp:330
NITER_SPA = 1000                                    # :329
MAX_EXP_LIM = 708                                   # src/Step2_Models.hpp:30
NL_DBL_DMIN = 10.0 * np.finfo(float).tiny           # src/Regenie.hpp:229


def _norm_cdf(x):
    return 0.5 * math.erfc(-x / math.sqrt(2.0))


def chisq1_from_pvalue(pv):
    """quantile(complement(chi2_1, p)) = (Phi^-1(1 - p/2))^2, by Newton on the erfc tail (src/Regenie.cpp:1859-1873)."""
    target = math.log(pv)
    z = math.sqrt(max(-2.0 * math.log(pv) - math.log(max(-2.0 * math.log(pv), 1.0)), 0.0)) if pv < 0.5 else 0.5
    for _ in range(200):
        f = math.erfc(z / math.sqrt(2.0))
        if f <= 0.0:
            z *= 0.5
            continue
        d = math.log(f) - target
        dz = d / (-math.sqrt(2.0 / math.pi) * math.exp(-0.5 * z * z) / f)
        z -= dz
        if abs(dz) < 1e-14 * max(1.0, abs(z)):
            break
    return z * z


def spa_test(stat, denum, gres, st, mask, nz, fast):
    """run_SPA_test_snp + solve_K1_snp + get_SPA_pvalue_snp (src/Step2_Models.cpp:2072-2294).
    gres: residualised genotype (length N); nz: g != 0 (the entries of Gsparse); fast = is_sparse.
    Returns (ok, chisq, logp)."""
    c = math.sqrt(denum)
    gmod = np.where(mask, gres / st.gamma_sqrt, 0.0)
    phat = st.phat
    gmu = gmod * phat
    a = gmu.sum()
    sel = (mask & nz) if fast else mask
    gm, ph, gs = gmod[sel], phat[sel], st.gamma_sqrt[sel]
    if fast:
        b = denum - float((gres[sel] ** 2).sum())
        d = float(gmu[sel].sum())
    lim_lo = gmod[gmod < 0].sum() - a
    lim_hi = gmod[gmod > 0].sum() - a
    score_num = stat * c
    if score_num < lim_lo or score_num > lim_hi:
        return False, 0.0, 0.0

    def K(t):
        v = np.log(1 - ph + ph * np.exp(t / c * gm)).sum()
        return v - t * d / c + t * t / 2 / denum * b if fast else v - t * a / c

    def K1(t):
        v = ((gm * ph / c) / (ph + (1 - ph) * np.exp(-t / c * gm))).sum()
        return v - d / c + t / denum * b if fast else v - a / c

    def K2(t):
        vexp = -t / c * gm
        if (vexp > MAX_EXP_LIM).any():
            return 0.0
        e = np.exp(vexp)
        v = ((gm * gm * gs * gs / (c * c) * e) / (ph + (1 - ph) * e) ** 2).sum()
        return v + b / denum if fast else v

    tval = -stat if stat >= 0 else stat
    ptot = 0.0
    for lam in (1, -1):
        if tval >= 0:
            min_x, max_x = 0.0, float(np.finfo(float).max)
        else:
            min_x, max_x = -float(np.finfo(float).max), 0.0
        t_old = 0.0
        f_old = lam * K1(lam * t_old) - tval
        t_new = -1.0
        it = 0
        while True:
            it += 1
            if it > NITER_SPA:
                return False, 0.0, 0.0
            hess = K2(lam * t_old)
            if hess == 0:
                return False, 0.0, 0.0
            t_new = t_old - f_old / hess
            f_new = lam * K1(lam * t_new) - tval
            if abs(f_new) < TOL_SPA:
                break
            if t_new and min_x < t_new < max_x:
                if f_new > 0:
                    max_x = t_new
                else:
                    min_x = t_new
            else:
                t_new = (min_x + max_x) / 2
                f_new = lam * K1(lam * t_new) - tval
                if f_new <= 0:
                    min_x = t_new
                else:
                    max_x = t_new
            t_old, f_old = t_new, f_new
        root = t_new
        kval = K(lam * root)
        k2 = K2(lam * root)
        if k2 == 0:
            return False, 0.0, 0.0
        wval = math.copysign(1.0, root) * math.sqrt(2 * (root * tval - kval)) if root != 0 else 0.0
        vval = root * math.sqrt(k2)
        if vval == 0:
            pv = 0.5
        else:
            pv = _norm_cdf(wval + math.log(vval / wval) / wval)
        ptot += pv
    if ptot > 1:
        return False, 0.0, 0.0
    pval = max(NL_DBL_DMIN, ptot)
    return True, chisq1_from_pvalue(pval), -math.log10(pval)


def score_bt(g_raw, info_term, in_analysis, mask, y_raw, st: BtChrom, z_thr, n_samples, male=None, non_par=False,
             correction="firth"):
    """One variant, one trait.  g_raw: dosages with -3 = missing.  Returns dict or None if ignored."""
    ok = in_analysis & (g_raw != -3.0)
    ns1 = int(ok.sum())
    total = float(g_raw[ok].sum())
    okp = ok & mask
    ns = int(okp.sum())
    tot_p = float(g_raw[okp].sum())
    info_p = float(info_term[okp].sum())
    if non_par and male is not None:                          # see step2.variant_stats
        mval = np.where(ok, g_raw, 0.0) * 0.5 * (2 - male.astype(float))
        m1, mp = float(mval.sum()), float(mval[mask].sum())
        mac1 = min(m1, 2 * ns1 - int((ok & male).sum()) - m1)
        mac = min(mp, 2 * ns - int((okp & male).sum()) - mp)
    else:
        mac1 = min(total, 2 * ns1 - total)
        mac = min(tot_p, 2 * ns - tot_p)
    if mac1 < MIN_MAC or mac < MIN_MAC:
        return None
    af = tot_p / (2.0 * ns)
    info = 1.0 if af in (0.0, 1.0) else 1 - info_p / (2 * ns * af * (1 - af))     # src/Geno.cpp:3140
    mean = total / ns1
    flipped = mean > 1                                                             # flip_geno :3150-3163
    g = g_raw.copy()
    if flipped:
        g = np.where(g != -3.0, 2 - g, g)
        mean = 2 - mean
    g = np.where(g == -3.0, mean, g)
    g = np.where(in_analysis, g, 0.0)
    is_sparse = int(((g != 0) & in_analysis).sum()) <= n_samples * (1 - PROP_ZERO_THR)
    gw = g * st.gamma_sqrt_mask
    if is_sparse:
        xtwg = st.Xg.T @ gw
        den = gw @ gw - xtwg @ xtwg
    else:
        gres = gw - st.Xg @ (st.Xg.T @ gw)
        den = gres @ gres
    sq = math.sqrt(den)
    if sq < NUMTOL:
        return None
    stat = (gw @ st.yres if is_sparse else gres @ st.yres) / sq
    out = dict(af=af, info=info, n=ns, flipped=flipped, stat=stat, test_fail=False)
    if abs(stat) <= z_thr:
        se = 1 / sq
        out.update(beta=stat * se, se=se, chisq=stat * stat, logp=get_logp(stat * stat))
    elif correction == "spa":
        if is_sparse:
            gres = gw - st.Xg @ xtwg
        ok_spa, chisq, logp = spa_test(stat, den, gres, st, mask, (g != 0), is_sparse)
        se0 = 1 / sq
        if not ok_spa:
            out.update(beta=stat * se0, se=se0, chisq=float("nan"), logp=float("nan"), test_fail=True)
        else:                                                                      # check_pval_snp :2021-2029
            out.update(beta=math.copysign(1.0, stat) * math.sqrt(chisq) * se0, se=se0, chisq=chisq, logp=logp)
    else:
        if is_sparse:
            gres = gw - st.Xg @ xtwg
        gvec = gres / st.gamma_sqrt                                               # :2056
        carriers = None
        if is_sparse and mac < 50:
            carriers = np.nonzero(mask & in_analysis & (g > 1e-4))[0]
        off = st.cov_blup_offset
        p = get_pvec(off)
        dev0 = logist_dev(y_raw, p, mask)
        if carriers is not None and len(carriers) > 0:
            pc = get_pvec(off[carriers])
            wc = np.where(mask[carriers], pc * (1 - pc), 1.0)
            dev0 -= math.log(((gvec[carriers] ** 2) * wc).sum())
            niter_pseudo = NITER_FIRTH // 2
        else:
            wv = np.where(mask, p * (1 - p), 1.0)
            dev0 -= math.log(((np.where(mask, gvec, 0.0) ** 2) * wv).sum())
            niter_pseudo = min(NITER_FIRTH // 2, 50)
        state, b, se, lrt = firth_pseudo_1snp(dev0, y_raw, gvec, off, mask, carriers, 0.0, niter_pseudo, NUMTOL_FIRTH)
        if state != 0:
            okf, b, se, lrt = firth_nr_1snp(dev0, y_raw, gvec, off, mask, carriers, 0.0, MAXSTEP, NITER_FIRTH // 2,
                                            NUMTOL_FIRTH)
            if not okf:
                se0 = 1 / sq
                out.update(beta=stat * se0, se=se0, chisq=float("nan"), logp=float("nan"), test_fail=True)
                out["beta"] = -out["beta"] if flipped else out["beta"]
                return out
        out.update(beta=b, se=se, chisq=lrt, logp=get_logp(lrt))
    if flipped:
        out["beta"] = -out["beta"]
    return out
