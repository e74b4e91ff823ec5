// Approximate Firth fallback for binary traits: one CTA per flagged (variant, trait).
// Replaces fit_firth_logistic_snp_fast -> fit_firth_pseudo / fit_firth for a single tested SNP with the
// covariate effects held in an offset (reference src/Step2_Models.cpp:1158-1252, 1527-1737), including the
// "carriers only" shortcut for sparse variants with MAC < 50.
//
// Every iteration of both solvers needs the same five sums over the active sample set S (masked samples, or
// the carriers): the deviance, X'WX = sum g^2 w, sum g p, sum g^3 w (1/2 - p) and a w == 0 flag; a pass
// evaluates them with one exp (+ one log) per sample and a fixed-order block reduction, so the scalar control
// flow of the reference runs uniformly in every thread of the CTA.
#include "kernels.cuh"

namespace rg {

constexpr int kFirthThreads = 512;
constexpr double kNumtolEps = 10.0 * 2.220446049250313e-16;

__device__ __forceinline__ double pvec(double eta) {   // get_pvec, src/Step1_Models.cpp:1797-1804
  if (eta > 30.0) return 1.0 / (1.0 + kNumtolEps);
  if (eta < -30.0) return kNumtolEps / (1.0 + kNumtolEps);
  return 1.0 - 1.0 / (exp(eta) + 1.0);
}

template <int K>
__device__ __forceinline__ void block_sum(double (&v)[K], double* sh) {
#pragma unroll
  for (int k = 0; k < K; ++k)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0)
#pragma unroll
    for (int k = 0; k < K; ++k) sh[warp * K + k] = v[k];
  __syncthreads();
#pragma unroll
  for (int k = 0; k < K; ++k) {
    double s = 0.0;
    for (int wq = 0; wq < kFirthThreads / 32; ++wq) s += sh[wq * K + k];
    v[k] = s;
  }
}

struct FirthSums { double dev, xtwx, gp, b3, w0; };

__global__ void __launch_bounds__(kFirthThreads)
s2_firth_kernel(FirthArgs a) {
  __shared__ double sh[(kFirthThreads / 32) * 6];
  const int sel = blockIdx.x;
  const int i = a.sel_var[sel], ph = a.sel_trait[sel];
  const int C = a.C, P = a.P;
  const int64_t npad = a.npad;
  const uint32_t* drow = a.dz + (int64_t)i * npad;
  const double* w = a.w + (int64_t)ph * npad;
  const double* gs = a.gs + (int64_t)ph * npad;
  const double* off = a.off + (int64_t)ph * npad;
  const double* xw = a.xw + (int64_t)ph * C * npad;
  const int8_t* ym = a.ym + (int64_t)ph * npad;
  double* gv = a.gvec + (int64_t)sel * npad;
  int8_t* cf = a.cflag + (int64_t)sel * npad;
  const int flags = a.flags[i];
  const bool flip = flags & 8, sparse = flags & 4;
  const double mu = a.mu[i];
  const bool try_fast = sparse && a.mac[(int64_t)i * P + ph] < 50.0;
  double v[kMaxCov];
  for (int c = 0; c < C; ++c) v[c] = a.xtwg[((int64_t)i * P + ph) * C + c];

  // ---- pass A: the residualised genotype G_res / Gamma^{1/2} and the carrier flags
  double cnt[1] = {0.0};
  for (int64_t t = threadIdx.x; t < npad; t += kFirthThreads) {
    const uint32_t dv = drow[t];
    double g = (dv & 0x80000000u) ? mu : (flip ? 2.0 - (double)(dv & 0x3FFu) / 255.0 : (double)(dv & 0x3FFu) / 255.0);
    if (a.F[t * a.dp] == 0.0) g = 0.0;
    double r = g * w[t];
    for (int c = 0; c < C; ++c) r -= xw[(int64_t)c * npad + t] * v[c];
    const double gsv = gs[t];
    gv[t] = (gsv != 0.0) ? r / gsv : 0.0;
    const bool car = try_fast && ym[t] != 0 && a.F[t * a.dp] != 0.0 && g > 1e-4;
    cf[t] = car ? 1 : 0;
    cnt[0] += car ? 1.0 : 0.0;
  }
  block_sum<1>(cnt, sh);
  const bool fast = try_fast && cnt[0] > 0.0;

  // sums over S at coefficient b; with_dev adds the deviance over S (and over all masked samples in dev_all)
  double sum_gy = 0.0;
  auto pass = [&](double b, bool with_dev, bool all_mask_dev, double* dev_all) {
    double s[6] = {0, 0, 0, 0, 0, 0};
    for (int64_t t = threadIdx.x; t < npad; t += kFirthThreads) {
      const int8_t code = ym[t];
      if (code == 0) continue;
      const bool inS = fast ? (cf[t] != 0) : true;
      if (!inS && !all_mask_dev) continue;
      const double g = gv[t];
      const double p = pvec(off[t] + g * b);
      if (with_dev) {
        const double ll = -2.0 * ((code == 1) ? log(1.0 - p) : log(p));
        s[5] += ll;
        if (inS) s[0] += ll;
      }
      if (inS) {
        const double wv = p * (1.0 - p), g2 = g * g;
        s[1] = fma(g2, wv, s[1]);
        s[2] = fma(g, p, s[2]);
        s[3] = fma(g2 * g, wv * (0.5 - p), s[3]);
        if (wv == 0.0) s[4] += 1.0;
      }
    }
    block_sum<6>(s, sh);
    if (dev_all) *dev_all = s[5];
    FirthSums r{s[0], s[1], s[2], s[3], s[4]};
    return r;
  };
  {  // sum g y over S (constant)
    double s[1] = {0.0};
    for (int64_t t = threadIdx.x; t < npad; t += kFirthThreads) {
      const int8_t code = ym[t];
      if (code == 2 && (!fast || cf[t])) s[0] += gv[t];
    }
    block_sum<1>(s, sh);
    sum_gy = s[0];
  }
  double dev_mask0;
  const FirthSums z = pass(0.0, true, true, &dev_mask0);
  const double dev0 = dev_mask0 - log(z.xtwx);
  const double dev_nc = fast ? dev_mask0 - z.dev : 0.0;
  const double tol = a.tol;
  const int niter_pseudo = fast ? a.niter / 2 : min(a.niter / 2, 50);

  // ---- fit_firth_pseudo (:1527-1641)
  int state = 1;
  double beta = 0.0, betanew = 0.0, b14 = 0.0, xtwx = 1.0, dev_new = 0.0;
  {
    int it = 0;
    bool converged = false;
    while (it < niter_pseudo) {
      ++it;
      const FirthSums r = (it == 1) ? z : pass(beta, true, false, nullptr);
      xtwx = r.xtwx;
      dev_new = dev_nc + r.dev - log(xtwx);
      const double ystar_g = sum_gy + r.b3 / xtwx;
      double score = ystar_g - r.gp;
      if (fabs(score) < tol && it >= 2) { converged = true; break; }
      if (it == 14) b14 = beta;
      if (it == 15 && fabs(beta - b14) > 0.1) { state = 1; goto pseudo_done; }
      int nl = 0;
      double bdiff = 1e16;
      while (nl < 25) {
        ++nl;
        const double step = score / xtwx, bnew = fabs(step);
        if (bnew > bdiff) { state = 2; goto pseudo_done; }
        const double mx = bnew / 5.0;
        betanew = beta + (mx > 1.0 ? step / mx : step);
        const FirthSums q = pass(betanew, false, false, nullptr);
        score = ystar_g - q.gp;
        if (fabs(score) < tol) break;
        if (q.w0 > 0.0) { state = 3; goto pseudo_done; }
        xtwx = q.xtwx;
        beta = betanew;
        bdiff = bnew;
      }
      beta = betanew;
    }
    if (converged) state = 0;
  }
pseudo_done:
  double se = 0.0, lrt = 0.0;
  int status = 0;
  if (state == 0) {
    lrt = dev0 - dev_new;
    if (lrt < 0.0) state = 4; else se = sqrt(1.0 / xtwx);
  }
  if (state != 0) {
    // ---- fit_firth, Newton-Raphson with step halving (:1644-1737)
    beta = 0.0;
    FirthSums r = z;
    xtwx = r.xtwx;
    double dev_old = dev_mask0 - log(xtwx);
    dev_new = dev_old;
    int it = 0;
    bool converged = false;
    const int niter = a.niter / 2;
    while (it < niter) {
      ++it;
      const double score = sum_gy - r.gp + r.b3 / xtwx;
      if (fabs(score) < tol && it >= 2) { converged = true; break; }
      double step = score / xtwx;
      const double mx = fabs(step) / a.maxstep;
      if (mx > 1.0) step /= mx;
      bool ok = false;
      for (int ls = 1; ls <= 25; ++ls) {
        if (ls > 1) step /= 2.0;
        r = pass(beta + step, true, false, nullptr);
        xtwx = r.xtwx;
        dev_new = dev_nc + r.dev - log(xtwx);
        if (dev_new < dev_old) { ok = true; break; }
      }
      if (!ok) step += 1e-6;
      beta += step;
      dev_old = dev_new;
    }
    if (converged) {
      lrt = dev0 - dev_new;
      if (lrt < 0.0) status = 1; else se = sqrt(1.0 / xtwx);
    } else {
      status = 1;
    }
    if (status == 0) status = 0;
  }
  if (threadIdx.x == 0) {
    a.beta[sel] = flip ? -beta : beta;
    a.se[sel] = se;
    a.lrt[sel] = lrt;
    a.status[sel] = status | (state << 4) | (fast ? 256 : 0);
  }
}

void launch_s2_firth(const FirthArgs& a, cudaStream_t s) {
  s2_firth_kernel<<<a.n_sel, kFirthThreads, 0, s>>>(a);
}

}  // namespace rg
