"""Step-2 per-variant score test, quantitative traits (oracle; test infrastructure only).

numpy-float64 restatement of rgcgithub/regenie v4.1.2:
  blup_read / blup_read_chr        src/Pheno.cpp:1241-1330, src/Step2_Models.cpp:51-143
  Data::compute_res                src/Data.cpp:2386-2404
  parseSnpfromBed                  src/Geno.cpp:2414-2536
  compute_mac / compute_aaf_info   src/Geno.cpp:3077-3148
  update_trait_counts              src/Geno.cpp:2948-2958
  check_sparse_G                   src/Geno.cpp:3165-3178
  residualize_geno                 src/Geno.cpp:3242-3262
  compute_score_qt                 src/Step2_Models.cpp:343-467
  get_logp                         src/Regenie.cpp:1843-1857
  print_sum_stats_head/_single     src/Step2_Models.cpp:2410-2417, 2502-2540

Parity status: QT Step 2 has no golden vector in the reference's tests ->
"parity unpinned" for compute_score_qt; the decode / AF / N / printing code is shared with
the binary-trait golden run.
"""
import math

import numpy as np

from .plink import MISSING_G
from .prep import MISSING, NUMTOL

MIN_MAC = 5.0          # src/Regenie.hpp:311
PROP_ZERO_THR = 0.5    # src/Regenie.hpp:311


def read_loco(path):
    """Parse a .loco file -> (ids, {chrom: row of strings})."""
    with open(path) as fh:
        hdr = fh.readline().split()
        if hdr[0] != "FID_IID":
            raise ValueError("header of blup file must start with FID_IID")
        rows = {}
        for line in fh:
            t = line.split()
            rows[int(t[0])] = t[1:]
    return hdr[1:], rows


def blup_mask(loco_ids, first_row, sample_index, n):
    """blup_read (src/Pheno.cpp:1283-1300): samples absent from the file or NA are masked."""
    m = np.zeros(n, dtype=bool)
    for k, v in zip(loco_ids, first_row):
        i = sample_index.get(k)
        if i is not None:
            m[i] = v != "NA"
    return m


def blup_chr(loco_ids, row, sample_index, n, in_analysis, mask_ph):
    """blup_read_chr (src/Step2_Models.cpp:96-124)."""
    b = np.zeros(n)
    for k, v in zip(loco_ids, row):
        i = sample_index.get(k)
        if i is None or not in_analysis[i] or not mask_ph[i]:
            continue
        if v == "NA":
            raise ValueError("individual has missing predictions (FID_IID=%s)" % k)
        b[i] = float(v)
    return b


def compute_res(Y, blups, mask, neff, ncov, scale_Y):
    """Data::compute_res (src/Data.cpp:2386-2404).  Returns res, p_sd_yres, scf_sv."""
    res = (Y - blups) * mask
    p_sd = np.linalg.norm(res, axis=0) / np.sqrt(neff - ncov)
    res = res / p_sd[None, :]
    return res, p_sd, scale_Y * p_sd


def get_logp(t):
    """src/Regenie.cpp:1843-1857; chi2_1 survival = erfc(sqrt(T/2))."""
    if t < 0 and abs(t) < 1e-6:
        return 0.0
    if t < 0:
        return -1.0
    pv = math.erfc(math.sqrt(t / 2.0))
    if pv == 0:
        lp = math.log10(2) - 0.5 * math.log10(2 * math.pi * t) - 0.5 * t * math.log10(math.e)
    else:
        lp = math.log10(pv)
    return -lp


def variant_stats(g_raw, in_analysis, mask, male=None, non_par=False):
    """parseSnpfromBed + compute_mac + compute_aaf_info for one variant (autosomal).

    g_raw: N hard calls with -3 = missing.  Returns dict with per-trait af, ns, mac, the
    all-trait af1/ns1/mac1, `ignored`, `ignored_trait` and the imputed genotype vector.
    """
    ok = in_analysis & (g_raw != MISSING_G)
    ns1 = int(ok.sum())
    total = float(g_raw[ok].sum())
    ns = (ok[:, None] & mask).sum(axis=0).astype(float)
    tot_p = (np.where(ok, g_raw, 0.0)[:, None] * mask).sum(axis=0)
    if non_par and male is not None:
        # non-PAR chrX: males (coded 0/2) count half; MAC = min(mac, 2N - N_males - mac)  (src/Geno.cpp:2447-2462, :3092-3097)
        mval = np.where(ok, g_raw, 0.0) * 0.5 * (2 - male.astype(float))
        macr1 = float(mval.sum())
        nmales1 = int((ok & male).sum())
        mac1 = min(macr1, 2 * ns1 - nmales1 - macr1)
        macr = (mval[:, None] * mask).sum(axis=0)
        nmales = ((ok & male)[:, None] & mask).sum(axis=0)
        mac = np.minimum(macr, 2 * ns - nmales - macr)
    else:
        mac1 = min(total, 2 * ns1 - total)
        mac = np.minimum(tot_p, 2 * ns - tot_p)
    out = dict(ns1=ns1, ns=ns.astype(int), mac1=mac1, mac=mac, ignored=mac1 < MIN_MAC,
               ignored_trait=mac < MIN_MAC)
    if out["ignored"]:
        return out
    out["af1"] = total / (2.0 * ns1)
    out["af"] = tot_p / (2.0 * ns)
    mean = total / ns1
    g = np.where(g_raw == MISSING_G, mean, g_raw)
    g = np.where(in_analysis, g, 0.0)                      # mean_impute_g src/Geno.cpp:3190-3193
    out["g"] = g
    return out


def score_qt(g, X, res, mask, in_analysis, n_analyzed, ncov, scf_sv, YtX, strict):
    """check_sparse_G + residualize_geno + compute_score_qt for one imputed variant.

    Returns dict(beta, se, chisq, logp, stats, is_sparse, scale_fac) or None when ignored.
    """
    n = len(g)
    is_sparse = int(((g != 0) & in_analysis).sum()) <= n * (1 - PROP_ZERO_THR)
    P = res.shape[1]
    if not is_sparse:
        g = g - X @ (X.T @ g)                               # src/Geno.cpp:3246-3247
        sf = np.linalg.norm(g) / math.sqrt(n_analyzed - ncov)
        if sf < NUMTOL:
            return None
        g = g / sf
        gsc = sf
        num = (res.T @ g) * gsc
        if strict:
            den = np.full(P, gsc * gsc * (n_analyzed - ncov))             # :387
        else:
            den = gsc * gsc * (mask.T.astype(float) @ (g * g))           # :416
    else:
        sf = 1.0
        gs = np.where(in_analysis, g, 0.0)
        XtG = X.T @ gs
        num = res.T @ gs - YtX @ XtG                        # :385 / :404
        if strict:
            den = np.full(P, gs @ gs - XtG @ XtG)           # :386
        else:
            den = np.empty(P)
            for ph in range(P):
                gm = gs * mask[:, ph]
                den[ph] = gm @ gm - 2 * (X.T @ gm) @ XtG + XtG @ XtG     # :410
    stats = num / np.sqrt(den)
    beta = stats * scf_sv / np.sqrt(den)
    se = beta / stats
    chisq = stats ** 2
    logp = np.array([get_logp(c) for c in chisq])
    return dict(beta=beta, se=se, chisq=chisq, logp=logp, stats=stats, is_sparse=is_sparse, scale_fac=sf,
                score=num, skat_var=den)               # dt_thr->scores / skat_var with --htp (:372-373, :391-394, :421-424)


def fmt(x):
    return "%g" % x


def sumstats_row(chrom, pos, vid, a0, a1, af, n, beta, se, chisq, logp, test="ADD", info=None,
                 test_pass=True):
    """print_sum_stats_head + print_sum_stats_single (native, split-by-phenotype format)."""
    s = "%d %d %s %s %s " % (chrom, pos, vid, a0, a1)
    s += (fmt(af) + " ") if af >= 0 else "NA "
    if info is not None:
        s += (fmt(info) + " ") if info >= 0 else "NA "
    s += "%d %s " % (n, test)
    if se >= 0 and not math.isnan(se):
        s += fmt(beta) + " " + fmt(se)
    else:
        s += "NA NA"
    if chisq >= 0 and test_pass and not math.isnan(logp):
        s += " " + fmt(chisq) + " " + fmt(logp)
    else:
        s += " NA NA"
    s += " " + ("NA" if test_pass else "TEST_FAIL") + "\n"
    return s


HEADER = "CHROM GENPOS ID ALLELE0 ALLELE1 A1FREQ N TEST BETA SE CHISQ LOG10P EXTRA\n"
HEADER_INFO = "CHROM GENPOS ID ALLELE0 ALLELE1 A1FREQ INFO N TEST BETA SE CHISQ LOG10P EXTRA\n"


# ------------------------------------------------------------------------------------------ --htp (HTPv4 rows)
ZCRIT = 1.959963984540054                      # quantile(complement(normal, .025)), src/Data.cpp:2118
LOG10_NL_DBL_DMIN = -math.log10(10.0 * 2.2250738585072014e-308)     # src/Regenie.hpp:229-230

HTP_HEADER = "\t".join(["Name", "Chr", "Pos", "Ref", "Alt", "Trait", "Cohort", "Model", "Effect", "LCI_Effect", "UCI_Effect",
                         "Pval", "AAF", "Num_Cases", "Cases_Ref", "Cases_Het", "Cases_Alt", "Num_Controls", "Controls_Ref",
                         "Controls_Het", "Controls_Alt", "Info"]) + "\n"      # print_header_output_htp, src/Step2_Models.cpp:2400


def convert_double_to_str(v):
    """src/Regenie.cpp:1691-1698."""
    return ("%.6f" % v) if (v < 5000 and v > 1e-5) else ("%g" % v)


def convert_logp_raw(logp, log_dbl_min=-math.log10(2.2250738585072014e-308) - 1):
    """src/Regenie.cpp:1700-1717 (default argument: src/Regenie.hpp:533)."""
    if logp <= 3:
        return "%f" % (10.0 ** -logp)
    if logp <= log_dbl_min:
        return "%g" % (10.0 ** -logp)
    thr = math.log(9.95) / math.log(10)
    base = math.ceil(logp)
    res = base - logp
    if res >= thr:
        res = 0
        base += 1
    return "%.1fe-%d" % (10.0 ** res, base)


def htp_model(test="ADD", wgr=True, bt=False, firth=False, spa=False):
    """test_string + wgr_string + correction_type, src/Data.cpp:2075-2102."""
    return test + ("-WGR" if wgr else "") + ("-FIRTH" if bt and firth else "-SPA" if bt and spa else "-LOG" if bt else "-LR")


def genocounts(g_raw, cases, controls=None):
    """update_genocounts (src/Geno.cpp:2986-3018) off chrX: thresholded counts (>= 1.5 alt, >= 0.5 het, < 0 missing) in the
    samples of `cases` (all samples with the trait for a QT) and, for a binary trait, of `controls`. -> 6 ints."""
    out = [0] * 6
    for k, idx in enumerate((cases, controls)):
        if idx is None:
            continue
        v = g_raw[idx]
        miss = int((v < 0).sum())
        alt = int((v >= 1.5).sum())
        het = int(((v >= 0.5) & (v < 1.5)).sum())
        out[3 * k: 3 * k + 3] = [len(idx) - het - alt - miss, het, alt]
    return out


def htp_row(vid, chrom, pos, a0, a1, trait, cohort, model, beta, se, chisq, lpv, af, mac, gc, test_pass=True, bt=False,
            firth=False, score=None, skat_var=None, cal_factor=-1.0, info=None):
    """print_sum_stats_head_htp + print_sum_stats_htp (src/Step2_Models.cpp:2419-2426, :2542-2646) for the single-variant
    tests of this repo (df = 1, no joint / burden columns).  gc = the 6 genotype counts of the trait."""
    s = "%s\t%d\t%d\t%s\t%s\t%s\t%s\t%s\t" % (vid, chrom, pos, a0, a1, trait, cohort, model)
    print_beta = test_pass and se >= 0 and not math.isnan(se)
    print_pv = test_pass and chisq >= 0 and not math.isnan(lpv)
    outp = "-1"
    if print_pv:
        if lpv > LOG10_NL_DBL_DMIN:
            outp = convert_logp_raw(LOG10_NL_DBL_DMIN)
        elif lpv > 0:
            outp = convert_logp_raw(lpv)
        else:
            outp = "0.9999999"
    outse = None
    if print_pv and not print_beta:
        s += "NA\tNA\tNA\t" + outp + "\t"
    elif not print_pv and not print_beta:
        s += "NA\tNA\tNA\tNA\t"
    elif (not bt) or (bt and firth and test_pass):
        if not bt:
            s += "%s\t%s\t%s\t" % (fmt(beta), fmt(beta - ZCRIT * se), fmt(beta + ZCRIT * se))
        else:
            s += "%s\t%s\t%s\t" % (fmt(math.exp(beta)), fmt(math.exp(beta - ZCRIT * se)), fmt(math.exp(beta + ZCRIT * se)))
        s += (outp if print_pv else "NA") + "\t"
    else:
        if print_pv:                                   # SPA / uncorrected logistic score test: allelic odds ratio
            eff = ((2 * gc[3] + gc[4] + .5) * (2 * gc[2] + gc[1] + .5) / (2 * gc[5] + gc[4] + .5) / (2 * gc[0] + gc[1] + .5))
            outse = abs(math.log(eff)) / math.sqrt(chisq)
            s += "%s\t%s\t%s\t%s\t" % (fmt(eff), fmt(eff * math.exp(-ZCRIT * outse)), fmt(eff * math.exp(ZCRIT * outse)), outp)
        else:
            s += "%s\t%s\t%s\tNA\t" % (fmt(math.exp(beta)), fmt(math.exp(beta - ZCRIT * se)), fmt(math.exp(beta + ZCRIT * se)))
    s += (fmt(af) + "\t") if af >= 0 else "NA\t"
    s += "%d\t%d\t%d\t%d\t" % (gc[0] + gc[1] + gc[2], gc[0], gc[1], gc[2])
    s += ("%d\t%d\t%d\t%d" % (gc[3] + gc[4] + gc[5], gc[3], gc[4], gc[5])) if bt else "NA\tNA\tNA\tNA"
    col = []
    if print_beta:
        if bt and test_pass:
            col += ["REGENIE_BETA=" + convert_double_to_str(beta), "REGENIE_SE=" + convert_double_to_str(se)]
            if print_pv and not firth:
                col.append("SE=" + convert_double_to_str(outse))
        elif bt:
            col += ["REGENIE_BETA=NA", "REGENIE_SE=NA"]
        else:
            col.append("REGENIE_SE=%f" % se)
    if info is not None and info >= 0:
        col.append("INFO=" + convert_double_to_str(info))
    if mac >= 0:
        col.append("MAC=%f" % mac)
    if score is not None:
        col.append("SCORE=" + convert_double_to_str(score))
    if skat_var is not None:
        col.append("SKATV=" + convert_double_to_str(skat_var * abs(cal_factor)))
    col.append("LOG10P=" + (convert_double_to_str(lpv) if print_pv else "NA"))
    if se < 0:
        col.append("NO_BETA")
    return s + "\t" + ";".join(col) + "\n"
