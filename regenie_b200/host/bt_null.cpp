#include "bt_null.hpp"

#include <algorithm>
#include <cmath>

namespace rgh {

namespace {

constexpr double kEps = 10.0 * 2.220446049250313e-16;   // numtol_eps, src/Regenie.hpp:225
constexpr double kNumtol = 1e-6;
constexpr int kNiterMax = 50, kNiterLs = 25, kNiterFirthNull = 1000, kMaxstepNull = 25;

inline double pvec(double eta) {   // get_pvec, src/Step1_Models.cpp:1797-1804
  if (eta > 30.0) return 1.0 / (1.0 + kEps);
  if (eta < -30.0) return kEps / (1.0 + kEps);
  return 1.0 - 1.0 / (std::exp(eta) + 1.0);
}

// in-place Cholesky of the SPD matrix a (n x n, row-major, lower); returns false if not positive definite
bool chol(std::vector<double>& a, int n) {
  for (int j = 0; j < n; ++j) {
    double d = a[(size_t)j * n + j];
    for (int k = 0; k < j; ++k) d -= a[(size_t)j * n + k] * a[(size_t)j * n + k];
    if (!(d > 0.0)) return false;
    d = std::sqrt(d);
    a[(size_t)j * n + j] = d;
    for (int i = j + 1; i < n; ++i) {
      double s = a[(size_t)i * n + j];
      for (int k = 0; k < j; ++k) s -= a[(size_t)i * n + k] * a[(size_t)j * n + k];
      a[(size_t)i * n + j] = s / d;
    }
  }
  return true;
}
void chol_solve(const std::vector<double>& l, int n, double* b) {
  for (int i = 0; i < n; ++i) {
    double s = b[i];
    for (int k = 0; k < i; ++k) s -= l[(size_t)i * n + k] * b[k];
    b[i] = s / l[(size_t)i * n + i];
  }
  for (int i = n - 1; i >= 0; --i) {
    double s = b[i];
    for (int k = i + 1; k < n; ++k) s -= l[(size_t)k * n + i] * b[k];
    b[i] = s / l[(size_t)i * n + i];
  }
}
double chol_logdet(const std::vector<double>& l, int n) {
  double s = 0.0;
  for (int i = 0; i < n; ++i) s += std::log(l[(size_t)i * n + i]);
  return 2.0 * s;
}

struct Work {
  int64_t N; int C;
  const double* y; const double* X; const uint8_t* mask;
  std::vector<double> eta, p;
  void eval(const double* offset, const double* beta) {
    for (int64_t i = 0; i < N; ++i) {
      double e = offset[i];
      for (int c = 0; c < C; ++c) e += X[(size_t)c * N + i] * beta[c];
      eta[i] = e;
      p[i] = pvec(e);
    }
  }
  double dev() const {   // get_logist_dev, src/Step1_Models.cpp:1819-1828
    double s = 0.0;
    for (int64_t i = 0; i < N; ++i)
      if (mask[i]) s += (y[i] == 0.0) ? std::log(1.0 - p[i]) : std::log(p[i]);
    return -2.0 * s;
  }
};

// IRLS with step halving (fit_logistic, src/Step1_Models.cpp:156-222)
bool fit_logistic(Work& w, const double* offset, std::vector<double>& beta, bool check_hs_dev) {
  const int64_t N = w.N; const int C = w.C;
  w.eval(offset, beta.data());
  double dev_old = w.dev(), dev_new = dev_old, diff_dev = 0.0;
  std::vector<double> betanew = beta, A((size_t)C * C), b(C);
  int it = 0;
  bool small_score = false, broke = false;
  while (it < kNiterMax) {
    ++it;
    std::fill(A.begin(), A.end(), 0.0);
    std::fill(b.begin(), b.end(), 0.0);
    for (int64_t i = 0; i < N; ++i) {
      if (!w.mask[i]) continue;
      const double wi = w.p[i] * (1.0 - w.p[i]);
      if (wi == 0.0) return false;
      const double z = w.eta[i] - offset[i] + (w.y[i] - w.p[i]) / wi;
      for (int c = 0; c < C; ++c) {
        const double xc = w.X[(size_t)c * N + i] * wi;
        b[c] += xc * z;
        for (int c2 = 0; c2 <= c; ++c2) A[(size_t)c * C + c2] += xc * w.X[(size_t)c2 * N + i];
      }
    }
    if (!chol(A, C)) return false;
    chol_solve(A, C, b.data());
    betanew = b;
    bool ok = false;
    for (int ls = 0; ls < kNiterLs; ++ls) {
      w.eval(offset, betanew.data());
      dev_new = w.dev();
      bool inside = true;
      for (int64_t i = 0; i < N && inside; ++i)
        if (w.mask[i] && !(w.p[i] > 0.0 && w.p[i] < 1.0)) inside = false;
      if (inside && (!check_hs_dev || dev_new < dev_old)) { ok = true; break; }
      for (int c = 0; c < C; ++c) betanew[c] = (beta[c] + betanew[c]) / 2.0;
    }
    if (!ok) return false;
    double smax = 0.0;
    for (int c = 0; c < C; ++c) {
      double s = 0.0;
      for (int64_t i = 0; i < N; ++i) if (w.mask[i]) s += w.X[(size_t)c * N + i] * (w.y[i] - w.p[i]);
      smax = std::max(smax, std::fabs(s));
    }
    if (smax < kNumtol) { broke = true; break; }
    if (!small_score && it < 20 && smax < 1.0) small_score = true;
    if (small_score && it > 20 && smax > 5.0) return false;
    diff_dev = std::fabs(dev_new - dev_old) / (0.1 + std::fabs(dev_new));
    beta = betanew;
    dev_old = dev_new;
  }
  if (!broke) ++it;
  if ((diff_dev == 0.0 || diff_dev >= kNumtol) && it > kNiterMax) return false;
  beta = betanew;
  return true;
}

// penalised-likelihood pieces at beta: deviance - log|X'WX|, modified score, X'WX factor
struct FirthEval { double dev; bool ok; };
FirthEval firth_eval(Work& w, const double* offset, const double* beta, std::vector<double>& L, std::vector<double>* score) {
  const int64_t N = w.N; const int C = w.C;
  w.eval(offset, beta);
  L.assign((size_t)C * C, 0.0);
  for (int64_t i = 0; i < N; ++i) {
    const double wi = w.mask[i] ? w.p[i] * (1.0 - w.p[i]) : 1.0;
    for (int c = 0; c < C; ++c) {
      const double xc = w.X[(size_t)c * N + i] * wi;
      if (xc == 0.0) continue;
      for (int c2 = 0; c2 <= c; ++c2) L[(size_t)c * C + c2] += xc * w.X[(size_t)c2 * N + i];
    }
  }
  if (!chol(L, C)) return {0.0, false};
  const double dev = w.dev() - chol_logdet(L, C);
  if (score) {
    score->assign(C, 0.0);
    std::vector<double> t(C);
    for (int64_t i = 0; i < N; ++i) {
      if (!w.mask[i]) continue;
      const double wi = w.p[i] * (1.0 - w.p[i]);
      // h_i = w_i x_i' (X'WX)^-1 x_i = w_i |L^-1 x_i|^2
      double h = 0.0;
      for (int c = 0; c < C; ++c) {
        double s = w.X[(size_t)c * N + i];
        for (int k = 0; k < c; ++k) s -= L[(size_t)c * C + k] * t[k];
        t[c] = s / L[(size_t)c * C + c];
        h += t[c] * t[c];
      }
      h *= wi;
      const double r = w.y[i] - w.p[i] + h * (0.5 - w.p[i]);
      for (int c = 0; c < C; ++c) (*score)[c] += w.X[(size_t)c * N + i] * r;
    }
  }
  return {dev, true};
}

// fit_firth_nr, all columns, null model (src/Step2_Models.cpp:1267-1383)
bool firth_nr(Work& w, const double* offset, std::vector<double>& beta, double maxstep, int niter, double tol) {
  const int C = w.C;
  double score_old = 1e16;
  int n_inc = 0, it = 0;
  std::vector<double> L, L2, score, step(C), bn(C);
  while (it < niter) {
    ++it;
    const FirthEval e = firth_eval(w, offset, beta.data(), L, &score);
    if (!e.ok) return false;
    const double dev_old = e.dev;
    step = score;
    chol_solve(L, C, step.data());
    double smax = 0.0, stepmax = 0.0;
    for (int c = 0; c < C; ++c) { smax = std::max(smax, std::fabs(score[c])); stepmax = std::max(stepmax, std::fabs(step[c])); }
    if (smax < tol && it >= 2) return true;
    n_inc = (smax > score_old) ? n_inc + 1 : 0;
    if (n_inc > 25) return false;
    const double mx = stepmax / maxstep;
    if (mx > 1.0) for (auto& s : step) s /= mx;
    bool ok = false;
    for (int ls = 1; ls <= kNiterLs; ++ls) {
      if (ls > 1) for (auto& s : step) s /= 2.0;
      for (int c = 0; c < C; ++c) bn[c] = beta[c] + step[c];
      const FirthEval e2 = firth_eval(w, offset, bn.data(), L2, nullptr);
      if (e2.ok && e2.dev < dev_old) { ok = true; break; }
    }
    if (!ok) return smax < tol;   // no decrease possible at a stationary point (a warm start that already is the solution)
    for (int c = 0; c < C; ++c) beta[c] += step[c];
    score_old = smax;
  }
  return false;
}

// the retry ladder of fit_approx_firth_null (src/Step2_Models.cpp:899-972) around the Newton solver: the given starting
// values, then zero, then zero with a fifth of the step and five times the iterations, then the given values again
bool firth_nr_ladder(Work& w, const double* offset, std::vector<double>& beta) {
  const std::vector<double> start = beta;
  double maxstep = kMaxstepNull;
  int niter = kNiterFirthNull;
  for (int trial = 0; trial < 4; ++trial) {
    std::vector<double> b = (trial == 0 || trial == 3) ? start : std::vector<double>(start.size(), 0.0);
    if (firth_nr(w, offset, b, maxstep, niter, 50 * kNumtol)) { beta = b; return true; }
    if (trial == 1) { maxstep /= 5; niter *= 5; }
  }
  return false;
}

}  // namespace

void get_basis(const std::vector<double>& X, int64_t N, int C0, std::vector<double>& Xb, int& nz);

BtNull fit_bt_null(const std::string& name, const double* y, const double* X, int64_t N, int C, const double* blup,
                   const uint8_t* mask, bool firth, const std::vector<double>* firth_start) {
  Work w{N, C, y, X, mask, std::vector<double>(N), std::vector<double>(N)};
  std::vector<double> loco(N);
  for (int64_t i = 0; i < N; ++i) loco[i] = blup[i] * (mask[i] ? 1.0 : 0.0);
  std::vector<double> beta(C, 0.0);
  bool ok = false;
  for (int chk = 1; chk >= 0 && !ok; --chk) {          // src/Step1_Models.cpp:79-86
    std::fill(beta.begin(), beta.end(), 0.0);
    ok = fit_logistic(w, loco.data(), beta, chk == 1);
  }
  if (!ok) throw Fail("logistic regression did not converge for phenotype '" + name + "'.");
  BtNull out;
  out.gamma_sqrt.resize(N); out.gamma_sqrt_mask.resize(N); out.yres.resize(N);
  out.y_hat_p = w.p;                                               // m_ests.Y_hat_p (src/Step1_Models.cpp:128)
  std::vector<double> XG((size_t)N * C);
  for (int64_t i = 0; i < N; ++i) {
    const double wi = mask[i] ? w.p[i] * (1.0 - w.p[i]) : 1.0;     // get_wvec, src/Step1_Models.cpp:1760
    const double g = std::sqrt(wi), gm = mask[i] ? g : 0.0;
    out.gamma_sqrt[i] = g;
    out.gamma_sqrt_mask[i] = gm;
    out.yres[i] = mask[i] ? (y[i] - w.p[i]) / g : 0.0;             // src/Data.cpp:2443-2445
    for (int c = 0; c < C; ++c) XG[(size_t)c * N + i] = gm * X[(size_t)c * N + i];
  }
  int nz = 0;
  get_basis(XG, N, C, out.x_gamma, nz);                            // src/Step1_Models.cpp:130-131
  if (nz != C) throw Fail("the weighted covariate matrix is rank deficient for phenotype '" + name + "'.");
  if (firth) {
    std::vector<double> bf = beta;
    if (firth_start) for (int c = 0; c < C && c < (int)firth_start->size(); ++c) bf[c] = (*firth_start)[c];   // get_beta_start_firth, :1936-1980
    if (!firth_nr_ladder(w, blup, bf))
      throw Fail("Firth penalized logistic regression failed to converge for phenotype '" + name + "' (null model).");
    out.firth_offset.resize(N);
    for (int64_t i = 0; i < N; ++i) {
      double e = blup[i];
      for (int c = 0; c < C; ++c) e += X[(size_t)c * N + i] * bf[c];
      out.firth_offset[i] = e;                                     // cov_blup_offset, src/Step2_Models.cpp:1014-1017
    }
  }
  return out;
}

std::vector<double> null_logistic_beta(const std::string& name, const double* y, const double* X, int64_t N, int C,
                                       const uint8_t* mask) {
  Work w{N, C, y, X, mask, std::vector<double>(N), std::vector<double>(N)};
  const std::vector<double> zero(N, 0.0);
  std::vector<double> beta(C, 0.0);
  bool ok = false;
  for (int chk = 1; chk >= 0 && !ok; --chk) {
    std::fill(beta.begin(), beta.end(), 0.0);
    ok = fit_logistic(w, zero.data(), beta, chk == 1);
  }
  if (!ok) throw Fail("logistic regression did not converge for phenotype '" + name + "'.");
  return beta;
}

bool fit_null_firth(const double* y, const double* X, int64_t N, int C, const double* blup, const uint8_t* mask,
                    std::vector<double>& beta) {
  Work w{N, C, y, X, mask, std::vector<double>(N), std::vector<double>(N)};
  std::vector<double> b = beta;
  b.resize(C, 0.0);
  if (!firth_nr_ladder(w, blup, b)) return false;
  beta = b;
  return true;
}

std::vector<double> null_logistic_eta(const std::string& name, const double* y, const double* X, int64_t N, int C,
                                      const uint8_t* mask) {
  Work w{N, C, y, X, mask, std::vector<double>(N), std::vector<double>(N)};
  const std::vector<double> zero(N, 0.0);
  std::vector<double> beta(C, 0.0);
  bool ok = false;
  for (int chk = 1; chk >= 0 && !ok; --chk) {
    std::fill(beta.begin(), beta.end(), 0.0);
    ok = fit_logistic(w, zero.data(), beta, chk == 1);
  }
  if (!ok) throw Fail("logistic regression did not converge for phenotype '" + name + "'.");
  return w.eta;
}

double chisq1_from_pvalue(double p) {
  if (!(p > 0.0)) return 0.0;
  if (p >= 1.0) return 0.0;
  // erfc(z / sqrt 2) = p, bisection on the (monotone) tail; 40 sigma covers 10 * DBL_MIN
  double lo = 0.0, hi = 40.0;
  for (int i = 0; i < 200; ++i) {
    const double mid = 0.5 * (lo + hi);
    if (std::erfc(mid / std::sqrt(2.0)) > p) lo = mid; else hi = mid;
  }
  const double z = 0.5 * (lo + hi);
  return z * z;
}

// inverse of the standard normal upper tail by bisection-safe Newton on erfc; chi2_1 quantile = z^2
double z_threshold(double p) {
  if (!(p > 0.0 && p < 1.0)) throw Fail("--pThresh must be in (0,1).");
  // solve erfc(z / sqrt 2) = p
  double lo = 0.0, hi = 40.0;
  for (int i = 0; i < 200; ++i) {
    const double mid = 0.5 * (lo + hi);
    if (std::erfc(mid / std::sqrt(2.0)) > p) lo = mid; else hi = mid;
  }
  return 0.5 * (lo + hi);
}

}  // namespace rgh
