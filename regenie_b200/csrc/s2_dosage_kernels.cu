// Step-2 statistics for 8-bit BGEN dosages and the binary-trait (logistic) score test.
// Replaces, per variant: parseSnpfromBGEN (reference src/Geno.cpp:2186-2345: dosage = p1/255 + 2 p0/255,
// INFO numerator 4 p0 + p1 - g^2, AF / N / MAC), flip_geno (:3150), mean imputation, check_sparse_G (:3165)
// and compute_score_bt (src/Step2_Models.cpp:470-556).
//
// As for hard calls (s2_kernels.cu) every statistic is a linear/quadratic form of code-wise sums of a
// per-sample feature row F; with dosages the "codes" are the integers d = p1 + 2 p0 (0..510) and
// e = 4 p0 + p1, so the kernel accumulates  S1 = sum d F,  S2 = sum d^2 F,  Sm = sum miss F,  Se = sum e F
// and the finish is closed form - including the minor-allele flip and the imputed mean, which are
// affine in d.  Feature row for binary traits (per trait p):
//   F = [ a | m_p | w_p^2 | w_p yres_p | w_p XW_pc (C) ]      w_p = Gamma_p^{1/2} m_p
#include "kernels.cuh"

namespace rg {

constexpr int kDzCols = 16;
constexpr int kDzSub = 128;

// BGEN probability rows [bs][n_file][2] (+ optional ploidy/missing bytes [bs][n_file]) -> padded sample
// layout, one uint32 per sample: d (bits 0-9) | e (bits 10-20) | missing (bit 31).  grid: (Npad/256, rows_p)
__global__ void dosage_relayout_kernel(const uint8_t* __restrict__ probs, const uint8_t* __restrict__ miss,
                                       int64_t n_file, int bs, const int32_t* __restrict__ file_idx_pad,
                                       int ref_first, uint32_t* __restrict__ dz, int64_t npad) {
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int row = blockIdx.y;
  if (t >= npad) return;
  uint32_t out = 0;
  const int fi = file_idx_pad[t];
  if (row < bs && fi >= 0) {
    const uint8_t* pr = probs + ((int64_t)row * n_file + fi) * 2;
    const uint32_t p0 = pr[0], p1 = pr[1];
    const bool m = miss ? (miss[(int64_t)row * n_file + fi] & 0x80) != 0 : false;
    if (m) {
      out = 0x80000000u;
    } else {
      // ref-last: g = p1 + 2 p0 (allele0 is ALT);  ref-first: g = p1 + 2 p2, p2 = 255 - p0 - p1 (>= 0)
      const uint32_t p2 = (p0 + p1 <= 255u) ? 255u - p0 - p1 : 0u;
      const uint32_t hom = ref_first ? p2 : p0;
      out = (p1 + 2u * hom) | ((4u * hom + p1) << 10);
    }
  }
  dz[(int64_t)row * npad + t] = out;
}

// grid: (rows_p/128, nchunks, Dp/16); block 128: thread = variant row.  part: [chunk][row][4][dp]
__global__ void __launch_bounds__(128)
dosage_stats_kernel(const uint32_t* __restrict__ dz, int64_t npad, const double* __restrict__ F, int dp,
                    const int4* __restrict__ chunks, int rows_p, double* __restrict__ part) {
  __shared__ double2 tile[kDzSub][kDzCols / 2];
  const int row = blockIdx.x * 128 + threadIdx.x;
  const int4 ch = chunks[blockIdx.y];
  const int col0 = blockIdx.z * kDzCols;
  const uint32_t* drow = dz + (int64_t)row * npad;
  double a1[kDzCols], a2[kDzCols], am[kDzCols], ae[kDzCols];
#pragma unroll
  for (int c = 0; c < kDzCols; ++c) a1[c] = a2[c] = am[c] = ae[c] = 0.0;
  for (int sub = 0; sub < ch.y; sub += kDzSub) {
    const int t0 = ch.x + sub;
    __syncthreads();
    for (int e = threadIdx.x; e < kDzSub * (kDzCols / 2); e += 128) {
      const int s = e / (kDzCols / 2), c2 = e % (kDzCols / 2);
      tile[s][c2] = *reinterpret_cast<const double2*>(F + (int64_t)(t0 + s) * dp + col0 + 2 * c2);
    }
    __syncthreads();
#pragma unroll 2
    for (int s = 0; s < kDzSub; ++s) {
      const uint32_t v = __ldg(drow + t0 + s);
      if (v == 0u) continue;
      const double2* xr = tile[s];
      if (v & 0x80000000u) {
#pragma unroll
        for (int c2 = 0; c2 < kDzCols / 2; ++c2) { const double2 f = xr[c2]; am[2 * c2] += f.x; am[2 * c2 + 1] += f.y; }
      } else {
        const double d = (double)(v & 0x3FFu), e = (double)((v >> 10) & 0x7FFu), d2 = d * d;
#pragma unroll
        for (int c2 = 0; c2 < kDzCols / 2; ++c2) {
          const double2 f = xr[c2];
          a1[2 * c2] = fma(d, f.x, a1[2 * c2]);   a1[2 * c2 + 1] = fma(d, f.y, a1[2 * c2 + 1]);
          a2[2 * c2] = fma(d2, f.x, a2[2 * c2]);  a2[2 * c2 + 1] = fma(d2, f.y, a2[2 * c2 + 1]);
          ae[2 * c2] = fma(e, f.x, ae[2 * c2]);   ae[2 * c2 + 1] = fma(e, f.y, ae[2 * c2 + 1]);
        }
      }
    }
  }
  double* o = part + (((int64_t)blockIdx.y * rows_p + row) * 4) * dp + col0;
#pragma unroll
  for (int c = 0; c < kDzCols; ++c) {
    o[c] = a1[c];
    o[dp + c] = a2[c];
    o[2 * dp + c] = am[c];
    o[3 * dp + c] = ae[c];
  }
}

// one thread per variant: AF / INFO / N / MAC, flip, sparse switch and the BT score test.
__global__ void s2_bt_finalize_kernel(S2BtFinalizeArgs a) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.bs) return;
  const int C = a.C, P = a.P, dp = a.dp;
  const double* S1 = a.sums + ((int64_t)i * 4) * dp;
  const double* S2 = S1 + dp;
  const double* Sm = S1 + 2 * dp;
  const double* Se = S1 + 3 * dp;
  const double k = 1.0 / a.unit, k2 = k * k;
  const double nm = Sm[0];
  const double ns1 = (double)a.n_analyzed - nm;
  const double total = S1[0] * k;                     // dosage sum over analysed, non-missing samples
  const bool xmale = a.non_par && a.non_par[i] && a.col_male >= 0;       // see s2_finalize_kernel
  double mac1;
  if (xmale) {
    const double macr = total - 0.5 * S1[a.col_male] * k;
    mac1 = fmin(macr, 2.0 * ns1 - (a.col_tot[a.col_male] - Sm[a.col_male]) - macr);
  } else {
    mac1 = fmin(total, 2.0 * ns1 - total);
  }
  int flags = 0;
  a.ns_all[i] = (int)ns1;
  a.mac_all[i] = mac1;
  a.af_all[i] = total / (2.0 * ns1);
  if (mac1 < a.min_mac) flags |= 1;
  double mu = total / ns1;
  const bool flip = a.with_flip && (mu > 1.0);        // flip_geno, src/Geno.cpp:3150-3163
  if (flip) { flags |= 8; mu = 2.0 - mu; }
  // non-zero entries among analysed samples after flip + imputation (check_sparse_G): a dosage is non-zero
  // iff d != 0, after a flip iff d != 510; imputed entries are non-zero iff the mean is.
  const double nnz = (flip ? ns1 - a.n510[i] : a.nz_count[i]) + ((mu != 0.0) ? nm : 0.0);
  const bool sparse = nnz <= (double)a.n_samples * 0.5;
  if (sparse) flags |= 4;
  a.scale_fac[i] = 1.0;
  a.flags[i] = flags;
  a.mu[i] = mu;
  for (int p = 0; p < P; ++p) {
    const int base = 1 + p * (3 + C);                 // [m_p | w^2 | w yres | w XW_c ...]
    const double nmiss_p = Sm[base];
    const double ns = a.col_tot[base] - nmiss_p;      // analysed & masked & non-missing
    const double tp = S1[base] * k;
    a.ns[(int64_t)i * P + p] = (int)ns;
    if (xmale) {
      const int cmale = a.col_male + 1 + p;
      const double macr = tp - 0.5 * S1[cmale] * k;
      a.mac[(int64_t)i * P + p] = fmin(macr, 2.0 * ns - (a.col_tot[cmale] - Sm[cmale]) - macr);
    } else {
      a.mac[(int64_t)i * P + p] = fmin(tp, 2.0 * ns - tp);
    }
    const double af = tp / (2.0 * ns);
    a.af[(int64_t)i * P + p] = af;
    // INFO (bgen), src/Geno.cpp:3140:  1 - sum(4 p0 + p1 - g^2) / (2 n af (1 - af))
    const double info_num = Se[base] * k - S2[base] * k2;
    a.info[(int64_t)i * P + p] = (af == 0.0 || af == 1.0) ? 1.0 : 1.0 - info_num / (2.0 * ns * af * (1.0 - af));
    // g_imp = (flip ? 2 - g : g) on non-missing, mu on missing.  For a feature f:
    //   sum g_imp f   = flip ? 2 (T - M) - s1 : s1,  + mu M      (T = sum f over analysed, M = sum over missing)
    //   sum g_imp^2 f = flip ? 4 (T - M) - 4 s1 + s2 : s2,  + mu^2 M
    auto lin = [&](int col) {
      const double s1 = S1[col] * k, M = Sm[col], T = a.col_tot[col];
      return (flip ? 2.0 * (T - M) - s1 : s1) + mu * M;
    };
    const int cw2 = base + 1, cwy = base + 2, cwx = base + 3;
    const double s1w2 = S1[cw2] * k, s2w2 = S2[cw2] * k2, Mw2 = Sm[cw2], Tw2 = a.col_tot[cw2];
    const double gw2 = (flip ? 4.0 * (Tw2 - Mw2) - 4.0 * s1w2 + s2w2 : s2w2) + mu * mu * Mw2;   // |GW|^2
    double xt2 = 0.0, xty = 0.0;
    for (int c = 0; c < C; ++c) {
      const double v = lin(cwx + c);                  // (XW^T GW)_c
      a.xtwg[((int64_t)i * P + p) * C + c] = v;
      xt2 += v * v;
      xty += v * a.xwy[(int64_t)p * C + c];           // (XW^T yres)_c
    }
    const double den = gw2 - xt2;                     // = |GW - XW XW^T GW|^2  (XW orthonormal)
    a.den[(int64_t)i * P + p] = den;
    double num = lin(cwy);                            // GW . yres
    if (!sparse) num -= xty;                          // dense path projects the covariates out of G first (:500,:521)
    const double sq = sqrt(den);
    double st = num / sq;
    if (!(sq >= a.numtol)) { st = 0.0; a.flags[i] |= 16; }
    const double se = 1.0 / sq;
    double beta = st * se;
    if (flip) beta = -beta;
    a.stat[(int64_t)i * P + p] = st;
    a.beta[(int64_t)i * P + p] = beta;
    a.se[(int64_t)i * P + p] = se;
    a.chisq[(int64_t)i * P + p] = st * st;
  }
}

// number of analysed samples with a non-zero dosage, per variant (for check_sparse_G). grid: rows_p, block 256
__global__ void dosage_nnz_kernel(const uint32_t* __restrict__ dz, int64_t npad, const double* __restrict__ F, int dp,
                                  double* __restrict__ out, double* __restrict__ out510) {
  __shared__ int red[256], red2[256];
  const uint32_t* drow = dz + (int64_t)blockIdx.x * npad;
  int c = 0, c2 = 0;
  for (int64_t t = threadIdx.x; t < npad; t += 256) {
    const uint32_t v = drow[t];
    if ((v & 0x80000000u) || F[t * dp] == 0.0) continue;
    c += ((v & 0x3FFu) != 0u) ? 1 : 0;
    c2 += ((v & 0x3FFu) == 510u) ? 1 : 0;
  }
  red[threadIdx.x] = c; red2[threadIdx.x] = c2;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) { red[threadIdx.x] += red[threadIdx.x + o]; red2[threadIdx.x] += red2[threadIdx.x + o]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) { out[blockIdx.x] = (double)red[0]; out510[blockIdx.x] = (double)red2[0]; }
}

// fixed-order sum over chunks.
__global__ void dosage_reduce_kernel(const double* __restrict__ part, int nchunks, int64_t per, double* __restrict__ sums) {
  const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= per) return;
  double s = 0.0;
  for (int c = 0; c < nchunks; ++c) s += part[(int64_t)c * per + e];
  sums[e] = s;
}

__global__ void dosage_scale_kernel(const double* __restrict__ s4, int rows_p, int dp, double* __restrict__ s3,
                                    double* __restrict__ se) {
  const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= (int64_t)rows_p * dp) return;
  const int64_t r = e / dp, c = e % dp;
  const double k = 1.0 / 255.0;
  const double* in = s4 + (r * 4) * dp + c;
  double* o = s3 + (r * 3) * dp + c;
  o[0] = in[0] * k;
  o[dp] = in[dp] * k * k;
  o[2 * dp] = in[2 * dp];
  se[r * dp + c] = in[3 * dp] * k;
}

void launch_dosage_scale(const double* sums4, int rows_p, int dp, double* sums3, double* se, cudaStream_t s) {
  dosage_scale_kernel<<<(unsigned)ceil_div((int64_t)rows_p * dp, 256), 256, 0, s>>>(sums4, rows_p, dp, sums3, se);
}

void launch_dosage_relayout(const uint8_t* probs, const uint8_t* miss, int64_t n_file, int bs, int rows_p,
                            const int32_t* file_idx_pad, int ref_first, uint32_t* dz, int64_t npad, cudaStream_t s) {
  dim3 grid((unsigned)ceil_div(npad, 256), rows_p);
  dosage_relayout_kernel<<<grid, 256, 0, s>>>(probs, miss, n_file, bs, file_idx_pad, ref_first, dz, npad);
}

void launch_dosage_stats(const uint32_t* dz, int64_t npad, const double* F, int dp, const int4* chunks, int nchunks,
                         int rows_p, double* part, double* sums, double* nnz, double* n510, cudaStream_t s) {
  dim3 grid(rows_p / 128, nchunks, dp / kDzCols);
  dosage_stats_kernel<<<grid, 128, 0, s>>>(dz, npad, F, dp, chunks, rows_p, part);
  const int64_t per = (int64_t)rows_p * 4 * dp;
  dosage_reduce_kernel<<<(unsigned)ceil_div(per, 256), 256, 0, s>>>(part, nchunks, per, sums);
  dosage_nnz_kernel<<<rows_p, 256, 0, s>>>(dz, npad, F, dp, nnz, n510);
}

void launch_s2_bt_finalize(const S2BtFinalizeArgs& a, cudaStream_t s) {
  s2_bt_finalize_kernel<<<(unsigned)ceil_div(a.bs, 128), 128, 0, s>>>(a);
}

}  // namespace rg
