// Per-chromosome null models of the binary-trait Step 2 (host side, O(N C^2) per trait and chromosome):
//   fit_null_logistic in test mode      reference src/Step1_Models.cpp:54-140, 156-222
//   Data::compute_res_bin               src/Data.cpp:2439-2455
//   fit_null_firth / fit_firth_nr       src/Step2_Models.cpp:985-1060, 1267-1383 (approximate Firth null)
// The per-variant work (score test, Firth fallback) runs on the GPU behind rg_s2_block_bgen8_bt / rg_s2_firth.
#pragma once
#include "data.hpp"

namespace rgh {

struct BtNull {
  std::vector<double> gamma_sqrt, gamma_sqrt_mask, yres, x_gamma, firth_offset, y_hat_p;   // [N] each, x_gamma [C][N]
};

// y (0/1), X [C][N] column-major orthonormal basis, blup (LOCO prediction), mask; throws Fail on non-convergence
// firth_start (optional, --use-null-firth): starting values of the null Firth fit instead of the logistic estimates
BtNull fit_bt_null(const std::string& name, const double* y, const double* X, int64_t N, int C, const double* blup,
                   const uint8_t* mask, bool firth, const std::vector<double>* firth_start = nullptr);

// --write-null-firth (Step 1, src/Data.cpp:1873-1903): coefficients of the covariate-only logistic fit (bhat_start) and
// one null approximate-Firth fit with the LOCO predictions of a chromosome as offset, warm-started from `beta`
// (in: starting values, out: estimates); false = did not converge
std::vector<double> null_logistic_beta(const std::string& name, const double* y, const double* X, int64_t N, int C,
                                       const uint8_t* mask);
bool fit_null_firth(const double* y, const double* X, int64_t N, int C, const double* blup, const uint8_t* mask,
                    std::vector<double>& beta);

// Step 1: linear predictor of the covariate-only logistic fit (offset_nullreg, fit_null_logistic called from
// src/Pheno.cpp:1608)
std::vector<double> null_logistic_eta(const std::string& name, const double* y, const double* X, int64_t N, int C,
                                      const uint8_t* mask);

// chi2_1 upper quantile of a p-value: quantile(complement(chisq1, p)) of get_logp (src/Regenie.cpp:1859-1873)
double chisq1_from_pvalue(double p);

// z threshold of --pThresh: sqrt of the chi2_1 upper quantile (src/Data.cpp:2116-2120)
double z_threshold(double p_thresh);

}  // namespace rgh
