// CPU restatement of regenie's Step-1 level-0 block and Step-2 per-variant score tests on Eigen 3.4.0 + OpenMP.
//
// TEST INFRASTRUCTURE ONLY (oracle/): the checker for tests/, the `cpu_baseline` leg and `--impl reference` arm of
// bench.py.  Nothing under regenie_b200/ links, loads or calls this.
//
// Why it exists: the reference binary cannot be built here (every src/*.cpp pulls Boost through Regenie.hpp and the
// BGEN library is not vendored), but its numerical kernels are plain Eigen expressions.  This file restates them with
// the same Eigen objects (MatrixXd products, SelfAdjointEigenSolver), the same loop structure and the same OpenMP
// placement, and is compiled against the reference's own vendored Eigen (external_libs/eigen-3.4.0) with the flags of
// the reference Makefile:33 (-O3 -ffast-math, -fopenmp on Linux :49).  It is therefore (a) an honest CPU baseline with
// the reference's threading behaviour (OpenMP over SNPs in the decoder, Eigen's OpenMP GEMM, a single-threaded
// eigensolver per fold) and (b) a second, Eigen-native pin for the numpy oracle (tests/test_ref_eigen_cpu.py).
//
// Restated functions (reference file:line):
//   readChunkFromBedFileToG   src/Geno.cpp:1702-1768   (+ buildLookupTable :2833-2857, mean_impute_g :3183-3188)
//   Data::residualize_genotypes   src/Data.cpp:190-228
//   Data::calc_cv_matrices (k-fold)   src/Data.cpp:729-776
//   ridge_level_0             src/Step1_Models.cpp:458-613
//   parseSnpfromBed / compute_mac / check_sparse_G / residualize_geno   src/Geno.cpp:2414-2536, 3077-3262
//   compute_score_qt          src/Step2_Models.cpp:343-467
//   compute_score_bt (score statistic only)   src/Step2_Models.cpp:470-556
#include <math.h>
#include <omp.h>
#include <stdint.h>

#include <chrono>
#include <vector>

#include <Eigen/Dense>

using Eigen::ArrayXd;
using Eigen::Map;
using Eigen::MatrixXd;
using Eigen::VectorXd;
typedef Eigen::Matrix<double, Eigen::Dynamic, Eigen::Dynamic, Eigen::ColMajor> MatCM;

namespace {

double now_s() {
  return std::chrono::duration<double>(std::chrono::high_resolution_clock::now().time_since_epoch()).count();
}

// 2-bit PLINK code -> allele count, ref-last: 00 -> 2, 01 -> missing (-3), 10 -> 1, 11 -> 0   (Geno.cpp:2843)
struct BedTable {
  double v[256][4];
  BedTable() {
    const double code[4] = {2, -3, 1, 0};
    for (int b = 0; b < 256; ++b)
      for (int j = 0; j < 4; ++j) v[b][j] = code[(b >> (2 * j)) & 3];
  }
};
const BedTable kTable;

void set_threads(int t) {
  if (t < 1) t = 1;
  omp_set_num_threads(t);
  Eigen::setNbThreads(t);
}

}  // namespace

extern "C" {

int rge_max_threads(void) { return omp_get_max_threads(); }
const char* rge_build_info(void) {
  return "Eigen "
#define RGE_STR2(x) #x
#define RGE_STR(x) RGE_STR2(x)
         RGE_STR(EIGEN_WORLD_VERSION) "." RGE_STR(EIGEN_MAJOR_VERSION) "." RGE_STR(EIGEN_MINOR_VERSION)
         " (reference external_libs), g++ -O3 -ffast-math -fopenmp"
#ifdef __AVX2__
         " -mavx2 -mfma"
#endif
      ;
}

// One Step-1 level-0 block, k-fold: packed .bed rows in, standardised level-0 predictors out.
//   bed        [bs][stride] PLINK rows (all N samples are kept: no --keep/--remove in the benchmark workloads)
//   X          [N x C] column-major orthonormal covariate basis (rows of samples outside the analysis are zero)
//   Y          [N x P] column-major residualised, scaled phenotypes (zero where masked)
//   mask       [N x P] column-major, 1 = phenotype observed
//   W_out      [P][N x R] column-major (the slab the reference writes with write_l0_file, Step1_Models.cpp:728)
//   phase_s    [4] seconds: decode+impute, residualise, cv matrices, ridge (eigensolver + predictions)
// Returns 0, or 1 + index of a SNP with sd below numtol (the reference throws there, Data.cpp:206-208).
int rge_l0_block_kfold(const uint8_t* bed, int64_t stride, int32_t bs, int64_t N, const uint8_t* in_analysis,
                       int32_t ref_first, const double* X, int32_t C, const double* Y, const uint8_t* mask, int32_t P,
                       const int64_t* fold_sizes, int32_t K, const double* lambda, int32_t R, const double* neff,
                       int64_t n_analyzed, int32_t threads, double* W_out, double* phase_s) {
  set_threads(threads);
  Map<const MatCM> Xm(X, N, C), Ym(Y, N, P);
  double t0 = now_s();

  // ---- readChunkFromBedFileToG: Gmat is bs x N column-major like the reference's MatrixXd; one OpenMP task per SNP
  MatrixXd G(bs, N);
  Eigen::setNbThreads(1);
#pragma omp parallel for schedule(dynamic)
  for (int j = 0; j < bs; ++j) {
    const uint8_t* row = bed + (int64_t)j * stride;
    double total = 0;
    int64_t ns = 0, i = 0;
    for (int64_t b = 0; b < stride; ++b) {
      const double* g4 = kTable.v[row[b]];
      for (int k = 0; k < 4 && i < N; ++k, ++i) {
        double hc = g4[k];
        if (ref_first && hc != -3) hc = 2 - hc;
        G(j, i) = hc;
        if (in_analysis[i] && hc != -3) { total += hc; ++ns; }
      }
    }
    total /= (double)ns;
    for (int64_t s = 0; s < N; ++s) {              // mean_impute_g
      double& g = G(j, s);
      if (!in_analysis[s]) g = 0;
      else if (g == -3) g = total;
    }
  }
  Eigen::setNbThreads(threads);
  double t1 = now_s();

  // ---- residualize_genotypes
  {
    ArrayXd keep(N);
    for (int64_t s = 0; s < N; ++s) keep(s) = in_analysis[s] ? 1.0 : 0.0;
    G.array().rowwise() *= keep.transpose();
  }
  MatrixXd beta_c = G * Xm;
  G -= beta_c * Xm.transpose();
  VectorXd scale_G = G.rowwise().norm() / sqrt((double)(n_analyzed - C));
  Eigen::Index min_i;
  if (scale_G.minCoeff(&min_i) < 1e-6) return 1 + (int)min_i;
  G.array().colwise() /= scale_G.array();
  double t2 = now_s();

  // ---- calc_cv_matrices, k-fold branch
  MatrixXd GGt = MatrixXd::Zero(bs, bs), GTY = MatrixXd::Zero(bs, P);
  std::vector<MatrixXd> G_folds(K), GtY(K);
  {
    int64_t off = 0;
    for (int f = 0; f < K; ++f) {
      Map<MatrixXd> Gf(&G(0, off), bs, fold_sizes[f]);
      GtY[f] = Gf * Ym.middleRows(off, fold_sizes[f]);
      GTY += GtY[f];
      G_folds[f] = Gf * Gf.transpose();
      GGt += G_folds[f];
      off += fold_sizes[f];
    }
  }
  double t3 = now_s();

  // ---- ridge_level_0: one eigendecomposition per fold, all ridge values from it, out-of-fold predictions
  MatrixXd p_sum = MatrixXd::Zero(R, P), p_sum2 = MatrixXd::Zero(R, P);
  {
    int64_t off = 0;
    for (int f = 0; f < K; ++f) {
      const int64_t nf = fold_sizes[f];
      MatrixXd ww1 = GGt - G_folds[f];
      Eigen::SelfAdjointEigenSolver<MatrixXd> eig(ww1);
      const MatrixXd& V = eig.eigenvectors();
      const VectorXd& d = eig.eigenvalues();
      MatrixXd ww2 = V.transpose() * (GTY - GtY[f]);
      MatrixXd mf(P, nf);
      for (int p = 0; p < P; ++p)
        for (int64_t s = 0; s < nf; ++s) mf(p, s) = mask[(int64_t)p * N + off + s] ? 1.0 : 0.0;
      for (int j = 0; j < R; ++j) {
        MatrixXd beta = V * (d.array() + lambda[j]).inverse().matrix().asDiagonal() * ww2;
        MatrixXd pred = ((beta.transpose() * G.block(0, off, bs, nf)).array() * mf.array()).matrix();   // P x nf
        p_sum.row(j) += pred.rowwise().sum().transpose();
        p_sum2.row(j) += pred.rowwise().squaredNorm().transpose();
        for (int p = 0; p < P; ++p) {
          double* w = W_out + ((int64_t)p * R + j) * N + off;
          for (int64_t s = 0; s < nf; ++s) w[s] = pred(p, s);
        }
      }
      off += nf;
    }
  }
  // centre and scale on the whole sample (Step1_Models.cpp:539-557)
  for (int p = 0; p < P; ++p)
    for (int j = 0; j < R; ++j) {
      const double mean = p_sum(j, p) / neff[p];
      const double invsd = sqrt((neff[p] - 1) / (p_sum2(j, p) - neff[p] * mean * mean));
      double* w = W_out + ((int64_t)p * R + j) * N;
      for (int64_t s = 0; s < N; ++s) w[s] = (w[s] - mean) * invsd;
    }
  double t4 = now_s();
  if (phase_s) { phase_s[0] = t1 - t0; phase_s[1] = t2 - t1; phase_s[2] = t3 - t2; phase_s[3] = t4 - t3; }
  return 0;
}

// Step 2, quantitative traits, .bed input: one OpenMP task per variant like Data::test_snps_fast (Data.cpp:2318-2346 ->
// parseSnpfromBed -> compute_score_qt), non-strict mode, additive test, no allele flip (QT).
//   res   [N x P] scaled residuals (compute_res), YtX [P x C], scf_sv [P]
//   out   per variant: af1, ns1, mac1, is_sparse, then per trait: af, ns, beta, se, chisq  -> [bs][4 + 5 P]
// A variant with MAC below min_mac gets af1 = -1 and no statistics.
int rge_s2_block_qt_bed(const uint8_t* bed, int64_t stride, int32_t bs, int64_t N, const uint8_t* in_analysis,
                        const double* X, int32_t C, const double* res, const uint8_t* mask, int32_t P,
                        const double* YtX, const double* scf_sv, int64_t n_analyzed, double min_mac, int32_t threads,
                        double* out) {
  set_threads(threads);
  Map<const MatCM> Xm(X, N, C), Rm(res, N, P), YtXm(YtX, P, C);
  MatrixXd maskd(N, P);
  for (int p = 0; p < P; ++p)
    for (int64_t s = 0; s < N; ++s) maskd(s, p) = mask[(int64_t)p * N + s] ? 1.0 : 0.0;
  const int64_t W = 4 + 5 * (int64_t)P;
  Eigen::setNbThreads(1);
#pragma omp parallel for schedule(dynamic)
  for (int v = 0; v < bs; ++v) {
    double* o = out + (int64_t)v * W;
    for (int64_t k = 0; k < W; ++k) o[k] = -1;
    const uint8_t* row = bed + (int64_t)v * stride;
    VectorXd g(N);
    ArrayXd tot_p = ArrayXd::Zero(P), ns_p = ArrayXd::Zero(P);
    double total = 0;
    int64_t ns1 = 0, i = 0;
    for (int64_t b = 0; b < stride; ++b) {
      const double* g4 = kTable.v[row[b]];
      for (int k = 0; k < 4 && i < N; ++k, ++i) {
        const double hc = g4[k];
        g(i) = hc;
        if (in_analysis[i] && hc != -3) {
          total += hc; ++ns1;
          for (int p = 0; p < P; ++p)
            if (mask[(int64_t)p * N + i]) { tot_p(p) += hc; ns_p(p) += 1; }
        }
      }
    }
    const double mac1 = std::min(total, 2.0 * ns1 - total);
    o[1] = (double)ns1; o[2] = mac1;
    if (mac1 < min_mac) continue;
    o[0] = total / (2.0 * ns1);
    const double mean = total / ns1;
    int64_t nnz = 0;
    for (int64_t s = 0; s < N; ++s) {
      if (!in_analysis[s]) g(s) = 0;
      else if (g(s) == -3) g(s) = mean;
      if (g(s) != 0) ++nnz;
    }
    const bool sparse = nnz <= N / 2;
    o[3] = sparse ? 1 : 0;
    ArrayXd num(P), den(P);
    if (!sparse) {                                     // residualize_geno + dense branch (Step2_Models.cpp:413-417)
      VectorXd b = Xm.transpose() * g;
      g -= Xm * b;
      const double sf = g.norm() / sqrt((double)(n_analyzed - C));
      if (sf < 1e-6) { o[0] = -1; continue; }
      g /= sf;
      num = (Rm.transpose() * g).array() * sf;
      den = sf * sf * (maskd.transpose() * g.array().square().matrix()).array();
    } else {                                           // sparse branch (Step2_Models.cpp:399-411)
      VectorXd XtG = Xm.transpose() * g;
      num = (Rm.transpose() * g - YtXm * XtG).array();
      const double ss = XtG.squaredNorm();
      for (int p = 0; p < P; ++p) {
        VectorXd gm = g.cwiseProduct(maskd.col(p));
        VectorXd XtGm = Xm.transpose() * gm;
        den(p) = gm.squaredNorm() - 2 * XtGm.dot(XtG) + ss;
      }
    }
    for (int p = 0; p < P; ++p) {
      const double st = num(p) / sqrt(den(p));
      const double bh = st * scf_sv[p] / sqrt(den(p));
      double* q = o + 4 + 5 * (int64_t)p;
      q[0] = tot_p(p) / (2.0 * ns_p(p)); q[1] = ns_p(p); q[2] = bh; q[3] = bh / st; q[4] = st * st;
    }
  }
  Eigen::setNbThreads(threads);
  return 0;
}

// Step 2, one binary trait, 8-bit BGEN probabilities already inflated: per sample (p0, p1) = P(hom first allele), P(het);
// default ref-last coding: g = p1/255 + 2 p0/255 (parseSnpfromBGEN fast path, Geno.cpp:2270-2279).
// Score statistic only (dense branch of compute_score_bt): stat = (G W - XG XG' G W)' yres / ||.||.
//   gsm [N] Gamma_sqrt * mask, XG [N x C] the orthonormal X_Gamma basis, yres [N]
//   out [bs][4]: af1, ns1, stat, denum
int rge_s2_block_bt_probs(const uint8_t* probs, const uint8_t* ploidy_missing, int32_t bs, int64_t N,
                          const uint8_t* in_analysis, const double* gsm, const double* XG, int32_t C,
                          const double* yres, double min_mac, int32_t threads, double* out) {
  set_threads(threads);
  Map<const MatCM> XGm(XG, N, C);
  Map<const VectorXd> gs(gsm, N), yr(yres, N);
  Eigen::setNbThreads(1);
#pragma omp parallel for schedule(dynamic)
  for (int v = 0; v < bs; ++v) {
    double* o = out + (int64_t)v * 4;
    o[0] = o[2] = o[3] = -1;
    const uint8_t* pr = probs + (int64_t)v * 2 * N;
    const uint8_t* pm = ploidy_missing + (int64_t)v * N;
    VectorXd g(N);
    double total = 0;
    int64_t ns = 0;
    for (int64_t s = 0; s < N; ++s) {
      if (pm[s] & 0x80) { g(s) = -3; continue; }
      const double d = (double)pr[2 * s + 1] / 255.0 + 2 * ((double)pr[2 * s] / 255.0);
      g(s) = d;
      if (in_analysis[s]) { total += d; ++ns; }
    }
    o[1] = (double)ns;
    const double mac = std::min(total, 2.0 * ns - total);
    if (mac < min_mac) continue;
    o[0] = total / (2.0 * ns);
    const bool flip = total / ns > 1.0;                 // minor-allele coding for non-QT additive tests (Data.cpp:2108)
    const double mean = total / ns;
    for (int64_t s = 0; s < N; ++s) {
      if (!in_analysis[s]) g(s) = 0;
      else { if (g(s) == -3) g(s) = mean; if (flip) g(s) = 2 - g(s); }
    }
    VectorXd GW = g.cwiseProduct(gs);
    VectorXd Gres = GW - XGm * (XGm.transpose() * GW);
    const double den = Gres.squaredNorm();
    o[3] = den;
    o[2] = Gres.dot(yr) / sqrt(den);
  }
  Eigen::setNbThreads(threads);
  return 0;
}

}  // extern "C"
