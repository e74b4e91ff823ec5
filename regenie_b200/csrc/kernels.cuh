// Kernel launchers and argument bundles shared between the .cu translation units.
#pragma once
#include "common.cuh"

namespace rg {

constexpr int kMaxFolds = 16;
constexpr int kMaxCov = 64;
constexpr int kMaxRidge = 8;
constexpr int kMaxPhenoTile = 8;
constexpr int kLimbs = 9;          // radix-30 digits per coefficient (44 bits)
constexpr int kLimbsI8 = 5;        // radix-254 int8 digits per coefficient of the INT8 prediction kernel (40 bits)
constexpr int kLimbQI8 = 50;       // outputs per INT8 prediction pass (5 x 50 = 250 <= 256 TMEM columns)
constexpr int kLimbQ = 56;         // outputs per tensor-core prediction pass (9 x 56 = 504 <= 512 TMEM columns)   // phenotypes per register pass of the LOOCV prediction kernel

// ---- bed_kernels.cu
void launch_bed_relayout(const uint8_t* packed, int64_t row_stride, int bs, int rows_p,
                         const int32_t* file_idx_pad, const int32_t* word_base, const uint32_t* word_keep, int ref_first,
                         uint32_t* gp, int64_t npad,
                         cudaStream_t s);
void launch_debug_sleep(unsigned ns, cudaStream_t s);
void launch_bed_expand_fp8(const uint32_t* gp, int rows_p, uint8_t* z, int64_t npad, cudaStream_t s);

// ---- l0_stats.cu
struct SnpFinalizeArgs {
  int bs, rows_p, C, P, K, cpp, loocv;
  long long n_analyzed;
  double numtol;
  const int32_t* cnt_fold;   // [K][rows_p][4]
  const double* sum_fold;    // [K][rows_p][2][cpp]
  const double* XtX_f;       // [K][C][C]
  const double* XtY_f;       // [K][C][P]
  double *mu, *inv_sd;       // [rows_p]
  double* Bv;                // [rows_p][C]
  double *Af, *Qf;           // [K][rows_p][C]
  double *gty_f, *rhs;       // [K][rows_p][P]
  unsigned long long* err_slot;
  long long err_base;
};

struct AssembleArgs {
  int bs, rows_p, nC, C, K, R, loocv;
  const float* zz;           // [K][2*rows_p][ldz] exact integer Grams
  int64_t ldz, zz_fold_stride;
  const double *mu, *inv_sd, *Bv, *Af, *Qf;
  const double* lambda;      // [R]
  double* cm;                // batched row-major lower systems
  int64_t cm_stride;
  int ldc;
  float* planes = nullptr;   // l0_assemble_sym only: FP32 hi / lo planes [K][2][n][n] of the same matrices (tensor-core operand)
  float* lplanes = nullptr;  // l0_assemble_sym only: factor planes [K*R][2][n][n] of the mixed solver; their first 128 columns
                             // receive A_f + lambda_r I (panel step 0 of the left-looking factorisation has nothing to subtract)
};

void launch_dbg_check_diag(const float* zz, int64_t ldz, int64_t fold_stride, const int32_t* cnt_fold, int rows_p,
                           int bs, int K, unsigned long long* counter, cudaStream_t s);
void launch_l0_stats(const uint32_t* gp, int64_t npad, const double* xy, int cpp, const int4* chunks,
                     int nchunks, int rows_p, int32_t* cnt_part, double* sum_part, cudaStream_t s);
void launch_l0_fold_reduce(const int32_t* cnt_part, const double* sum_part, int rows_p, int cpp,
                           const int2* fold_chunks, int K, int32_t* cnt_fold, double* sum_fold,
                           cudaStream_t s);
void launch_l0_snp_finalize(const SnpFinalizeArgs& a, cudaStream_t s);
void launch_l0_assemble(const AssembleArgs& a, const double* rhs, int P, int Ppad, int nmat, cudaStream_t s);

// ---- gram_tcgen05.cu
void make_gram_tensor_map(CUtensorMap* tm, const uint8_t* z, int64_t npad, int rows2);
size_t gram_smem_bytes(int bn = 256);
// ---- l0_stats_tc.cu: the statistics as extra Gram column tiles
constexpr int kStatQ = 14;          // xy columns per 128-row digit group (14 x 9 limbs = 126 rows)
constexpr int kStatOnesRow = 126;   // row of the all-ones column (group 0)
void launch_l0_xy_digits(const double* xy, int cpp, int ncol, int64_t npad, const uint8_t* is_real, double* scale,
                         uint8_t* D, cudaStream_t s);
void launch_l0_stats_finish(const float* T, int ldt, int64_t t_fold_stride, const float* zz, int ldz,
                            int64_t zz_fold_stride, int rows_p, int cpp, int ncol, int K, const double* scale,
                            int32_t* cnt_fold, double* sum_fold, cudaStream_t s);
void gram_tile_list(int rows2, std::vector<int2>& tiles);
void launch_gram_tcgen05(const CUtensorMap& tm, const CUtensorMap& tmB, const int2* tiles, int ntiles, const int2* fold_k, int K,
                         float* out, int ldo, int64_t fold_stride, float out_scale, cudaStream_t s, int bn = 256);
// operand-plane bytes of the Step-1 block (bed_expand_fp8_kernel): dosage d -> 8 d as int8 = 2^-6 d as e4m3
constexpr float kZScaleGram = 4096.f;     // Z Z^T tiles: both operands carry 2^-6
constexpr float kZScaleStat = 64.f;       // Z [X|Y]-digit tiles: the digit rows are plain e4m3 integers
void launch_gram_reference(const uint8_t* z, int64_t npad, int rows2, int k0, int k1, float* out, int ldo,
                           cudaStream_t s);

// ---- chol.cu
void launch_chol_factor(double* cm, int64_t stride, int nC, int n_aug, int batch, double* inv,
                        unsigned long long* err_slot, long long err_base, cudaStream_t s);
void launch_chol_backsolve(double* cm, int64_t stride, int nC, int P, int batch, const double* inv,
                           cudaStream_t s);
int chol_num_launches(int nC);
size_t chol_inv_elems(int nC, int batch);
void launch_chol_rows_backsolve(double* cm, int64_t stride, int nC, int row0, int nrows, int batch,
                                const double* inv, cudaStream_t s);

// ---- tf32_gemm.cu: batched 128x128 "NT" tiles in 3xTF32 on tcgen05 (operands = hi/lo FP32 planes [batch][2][n][n])
struct Tf32GemmEpilogue {
  int n;                      // matrix dimension = row stride of every output
  int64_t out_mat_stride;     // elements per plane per matrix (n * n)
  float* out;                 // D   as hi / lo planes [batch][2][n][n], or null
  float* out_t;               // D^T as hi / lo planes, or null
  float* out_plain;           // D   as one FP32 plane [batch][n][n], or null
  int mirror;                 // out_plain: also store D^T (symmetric result computed on its lower tiles)
  int negate;                 // D = -acc
  int lower_only;             // zero the strict upper part of diagonal tiles
  const double* cin;          // D = (float)(cin - acc) with FP64 matrices cin[mat / cin_mat_div][row][col], or null
  int64_t cin_mat_stride;
  int cin_ld, cin_mat_div;
  const double* diag_add;     // added to the diagonal of diagonal tiles: diag_add[mat % diag_mod] (the ridge shift), or null
  int diag_mod;
  int c_chunks;               // 0, or 4: leading K chunks  acc = C_tile * I  followed by the main chunks with A negated
  int c_mat_div;              // C matrix index = mat / c_mat_div
  int l2_prefetch;            // K chunks whose operand boxes are prefetched into L2 ahead of the single shared-memory stage
};
void make_tf32_planes_tensor_map(CUtensorMap* tm, const float* planes, int n, int batch);
void make_tf32_identity_planes(DevBuf<float>& buf, CUtensorMap* tm);
void make_f32_rows_tensor_map(CUtensorMap* tm, const float* base, int cols, int64_t rows, int box_cols, int box_rows);
void launch_tf32x3_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const int4* tiles, int ntiles, int batch,
                        const Tf32GemmEpilogue& ep, cudaStream_t s, const CUtensorMap* tmC = nullptr,
                        const CUtensorMap* tmI = nullptr);

// ---- chol_mixed.cu: tensor-core factorisation + FP64 iterative refinement of the level-0 ridge systems
constexpr int kMxMaxSteps = 6;
class MixedSolver {
 public:
  MixedSolver();
  ~MixedSolver();
  MixedSolver(const MixedSolver&) = delete;
  MixedSolver& operator=(const MixedSolver&) = delete;
  static int dim_for(int bs);                 // 128 * 2^k >= bs, or 0 when the block is too large for this path
  void prepare(int n, int K, int R, int Pp);
  // Af [K][n][n] FP64 full symmetric (no ridge shift) and its FP32 hi / lo planes Ap [K][2][n][n], lambda [R],
  // bvec [K][Pp][n]; xvec / rvec [K*R][Pp][n]
  float* a_planes();                          // where the assembler writes Ap (owned by the solver, valid after prepare)
  float* l_planes();                          // factor planes: the assembler may fill block column 0 (first_col_ready)
  void solve(const double* Af, const double* lambda, const double* bvec, double* xvec, double* rvec, int P, int steps,
             float tol, unsigned int* fail_flag, cudaStream_t s, bool first_col_ready = false);
  static int launches_per_solve(int n, int steps, int P);
  const float* debug_planes(int which) const;  // 0 L, 1 W, 2 W^T (hi/lo planes), 3 X
 private:
  struct Impl;
  Impl* impl;
};
// K full symmetric fold systems (no ridge shift) + right-hand sides as rows, for the mixed solver
void launch_l0_assemble_sym(const AssembleArgs& a, const double* rhs, int P, int Pp, double* bvec, cudaStream_t s);

// ---- l0_dense.cu: level-0 block on real-valued genotypes (8-bit dosages / FP64), dense FP64 like the reference
void launch_dense_from_dosage(const uint8_t* probs, const uint8_t* miss, int64_t n_file, int bs, const int32_t* file_idx_pad,
                              int ref_first, double* gd, int64_t npad, cudaStream_t s);
void launch_dense_from_f64(const double* G, int64_t n_file, int bs, const int32_t* file_idx_pad, double* gd, int64_t npad,
                           cudaStream_t s);
void launch_dense_prepare(double* gd, int64_t npad, int bs, const int32_t* file_idx_pad, const double* xy, int cpp, int C,
                          long long n_analyzed, double numtol, double* mu, double* sd, unsigned long long* err_slot,
                          long long err_base, cudaStream_t s);
void launch_dense_assemble(const double* part, int64_t part_stride, int ldp, const double* part_y, int64_t part_y_stride,
                           const int2* fold_chunks, int K, int R, const double* lambda, int bs, int nC, int P, double* cm,
                           int64_t cm_stride, int loocv, cudaStream_t s);
void launch_dense_loocv_fill(const double* gd, int64_t npad, int bs, int nC, double* cm, int64_t cm_stride, int row0, int R,
                             cudaStream_t s);
void launch_dense_predict(const double* gd, int64_t npad, int bs, const double* cm, int64_t cm_stride, int ldc, int nC, int R,
                          int P, const int32_t* tile_fold, const uint8_t* mask, double* const* W, int col0, cudaStream_t s);

// ---- l0_predict.cu
struct PredictArgs {
  int bs, rows_p, C, P, R, Qp, cpp, col0;
  int64_t npad, words_per_row;
  const uint32_t* gp;
  const int32_t* tile_fold;  // [npad/128]
  const double *gam, *gmu;   // [K][rows_p][Qp]
  const double* cvec;        // [K][Qp][C]
  const double* xy;          // [npad][cpp]
  const uint8_t* mask;       // [P][npad]
  double* const* W;          // [P] base of each phenotype's npad x B column-major predictor matrix (may be peer memory)
  double* part;              // [ntiles][Qp][2]
};
void launch_l0_gamma(const double* cm, int64_t cm_stride, int ldc, int nC, int R, int P, int Qp,
                     int bs, int rows_p, int K, const double* mu, const double* inv_sd,
                     const double* Bv, int C, double* gam, double* gmu, double* cvec, cudaStream_t s);
void launch_l0_predict(const PredictArgs& a, int ntiles, cudaStream_t s);
void launch_l0_standardize(const double* part, int ntiles, int Qp, int Q, int P, const double* neff,
                           double* mean_invsd, double* const* W, int64_t npad, int col0,
                           const uint8_t* is_real, cudaStream_t s, const double* const* src = nullptr, int src_col0 = 0);
int predict_qt();

// ---- predict_tcgen05.cu
struct PredictTcArgs {
  int rows_p, C, P, Q, Qp, cpp, col0, ngroups;
  int64_t npad;
  const int32_t* tile_fold;
  const double* scale;       // [K][Qp]
  const double* cvec;        // [K][Qp][C]
  const double* xy;
  const uint8_t* mask;
  double* const* W;
  double* part;
  long long* dbg;            // optional per-CTA clock64 stamps (profiling aid)
  int l2_prefetch = 0;       // INT8 kernel: k-blocks of the genotype planes prefetched into L2 ahead of the ring
};
void make_byte_tensor_map(CUtensorMap* tm, const uint8_t* basep, int64_t inner, int64_t rows);
size_t predict_tc_dig_bytes(int K, int ngroups, int rows_p);
void launch_l0_gamma_limbs(const double* gam, const double* gmu, int Qp, int Q, int bs, int rows_p, int K,
                           double* scale, uint8_t* dig, int ngroups, cudaStream_t s);
int launch_l0_colsum(double* const* W, int64_t npad, int col0, int P, int Q, int Qp, double* part,
                     cudaStream_t s);
void launch_l0_predict_tcgen05(const CUtensorMap& tmZ, const CUtensorMap& tmD, const PredictTcArgs& a, int ntiles,
                               cudaStream_t s);
// INT8 variant (kind::i8): 5 radix-254 digit rows per output, 256 TMEM columns, two CTAs per SM
size_t predict_i8_dig_bytes(int K, int ngroups, int rows_p);
void launch_l0_gamma_limbs_i8(const double* gam, const double* gmu, int Qp, int Q, int bs, int rows_p, int K,
                              double* scale, uint8_t* dig, int ngroups, cudaStream_t s);
void launch_l0_predict_i8(const CUtensorMap& tmZ, const CUtensorMap& tmD, const PredictTcArgs& a, int ntiles,
                          cudaStream_t s);

// ---- l1_kernels.cu
void launch_l1_gram(const double* W, int64_t ldw, int B, const int4* chunks, int nchunks, double* part,
                    int64_t part_stride, int ldp, cudaStream_t s);
void launch_l1_xty(const double* W, int64_t ldw, const double* xy, int cpp, int ycol, const int4* chunks,
                   int nchunks, double* part_y, int B, cudaStream_t s);
void launch_l1_assemble(const double* part, int64_t part_stride, int ldp, const double* part_y,
                        const int2* fold_chunks, int K, int R1, const double* tau, int B, int nC, double* cm,
                        int64_t cm_stride, int loocv, cudaStream_t s);
void launch_l1_pred_sums(const double* W, int64_t ldw, int B, int R1, const double* beta, int ldb,
                         const int32_t* tile_fold, const double* xy, int cpp, int ycol, double* part_out,
                         int ntiles, double* out, cudaStream_t s);
void launch_l1_chr_pred(const double* W, int64_t ldw, int nchr, const int32_t* chr_col_start, const double* beta,
                        int ldb, int R1, int best, const int32_t* tile_fold, double* pred, int64_t npad,
                        cudaStream_t s);

// ---- loocv_kernels.cu
void launch_l0_loocv_fill(const uint32_t* gp, int64_t npad, int bs, int nC, const double* mu, const double* inv_sd,
                          const double* Bv, int C, const double* xy, int cpp, double* cm, int64_t cm_stride,
                          int nrow0, int R, cudaStream_t s);
void launch_l0_loocv_pred(const double* cm, int64_t cm_stride, int nC, int bs, int Ppad, int P, int R,
                          const double* xy, int cpp, int C, const uint8_t* mask, int64_t npad, double* const* W,
                          int col0, double* part, int Qp, cudaStream_t s);
void launch_l0_loocv_std_apply(double* const* W, int64_t npad, int col0, int P, int Q, const uint8_t* mask,
                               const double* mean_invsd, cudaStream_t s);
void launch_l1_loocv_fill(const double* W, int64_t ldw, int B, int nC, double* cm, int64_t cm_stride, int nrow0,
                          int R1, int64_t npad, cudaStream_t s);
void launch_l1_loocv_sums(const double* cm, int64_t cm_stride, int nC, int B, int nrow0, const double* xy, int cpp,
                          int ycol, double* part, int R1, int ntiles, double* out, cudaStream_t s);
void launch_rows_sqnorm(const double* rows, int nC, int B, double* out, int ntiles, cudaStream_t s);

// ---- l1_logistic.cu
void launch_l1_scale_rows(const double* W, int64_t ldw, int B, const double* wm, double* Ws, cudaStream_t s);
void launch_l1_bt_eta(const double* W, int64_t ldw, int B, const double* beta, const double* offset, const int8_t* ym,
                      double* eta, double* pv, double* wm, double* resid, double* dev_part, double* dev_out,
                      cudaStream_t s);
void launch_l1_bt_score(const double* part_y, int nchunks, int B, int nC, double tau, const double* beta, double* score,
                        double* rhs_row, cudaStream_t s);
void launch_l1_bt_loo_sums(const double* eta, const double* q, const double* wm, const double* resid, const int8_t* ym,
                           double eps, double* fvec, double* part, double* out6, int64_t npad, cudaStream_t s);
void launch_l1_bt_chr_pred(const double* W, int64_t ldw, int nC, const double* zrows, const double* fvec,
                           const double* bvec, int nchr, const int32_t* chr_col_start, double* pred, int64_t npad,
                           cudaStream_t s);
void launch_l1_loocv_chr_pred(const double* W, int64_t ldw, int B, int nC, const double* zrows, const double* hvec,
                              const double* bvec, const double* xy, int cpp, int ycol, int nchr,
                              const int32_t* chr_col_start, double* pred, int64_t npad, cudaStream_t s);
void launch_l0_std_reduce_only(const double* part, int ntiles, int Qp, int Q, int P, const double* neff,
                               double* mean_invsd, cudaStream_t s);

// ---- s2_kernels.cu
struct S2FinalizeArgs {
  int bs, C, P, dp, strict;
  long long n_analyzed, n_samples;
  double min_mac, numtol;
  const double* sums;        // [rows_p][3][dp]
  const double* mask_count;  // [P]
  const double* YtX;         // [P][C]
  const double* XmX;         // [P][C][C]
  const double* scf_sv;      // [P]
  const uint8_t* non_par = nullptr;   // [bs] variant lies in the non-PAR part of chrX (males count half towards MAC)
  int col_male = -1;                  // F column of the male indicator (followed by male x mask_p), -1 = no sex information
  const double* male_tot = nullptr;   // [1 + P] analysed males, masked-in males per trait
  const double* nz_count = nullptr;   // [rows_p] non-zero dosages among analysed samples (dosage input only)
  const double* info_sums = nullptr;  // [rows_p][dp] sum (4 p0 + p1) F (dosage input only)
  double* info = nullptr;             // [bs x P] INFO (dosage input only)
  double *af, *mac, *af_all, *mac_all, *scale_fac, *stat, *beta, *se, *chisq;
  int32_t *ns, *ns_all, *flags;
};
void launch_s2_stats(const uint32_t* gp, int64_t npad, const double* F, int dp, const int4* chunks, int nchunks,
                     int rows_p, double* part, double* sums, cudaStream_t s);
void launch_s2_finalize(const S2FinalizeArgs& a, cudaStream_t s);
void launch_bed_expand3_fp8(const uint32_t* gp, int rows_p, uint8_t* z, int64_t npad, cudaStream_t s);
// hard calls for the binary-trait path: tensor sums -> [rows][4][dp] (S1, S2, Sm, 0) + non-zero / hom-alt counts + dz words
void launch_s2_bt_bed_finish(const float* T, int ldt, int64_t chunk_stride, int nchunk, int rows_p, int dp, int D,
                             const double* scale, double* sums4, double* nnz, double* n2, cudaStream_t s);
void launch_gp_to_dz(const uint32_t* gp, int rows_p, uint32_t* dz, int64_t npad, cudaStream_t s);
void launch_s2_stats_finish(const float* T, int ldt, int64_t chunk_stride, int nchunk, int rows_p, int dp, int D,
                            const double* scale, double* sums, cudaStream_t s);

// ---- s2_dosage_kernels.cu
struct S2BtFinalizeArgs {
  int bs, C, P, dp, with_flip;
  long long n_analyzed, n_samples;
  double min_mac, numtol;
  const double* sums;        // [rows_p][4][dp]  S1, S2, Sm, Se in integer dosage units
  const double* col_tot;     // [dp] sum of every feature column over all samples
  const double* xwy;         // [P][C]  XW^T yres
  const uint8_t* non_par = nullptr;   // see S2FinalizeArgs
  int col_male = -1;
  double unit = 255.0;                // S1 / S2 / Se are in units of 1/unit (255 for 8-bit dosages, 1 for hard calls)
  const double* nz_count;    // [rows_p] analysed samples with non-zero dosage
  const double* n510;        // [rows_p] analysed samples with dosage exactly 2
  double *af, *mac, *info, *af_all, *mac_all, *scale_fac, *stat, *beta, *se, *chisq, *xtwg, *mu, *den;
  int32_t *ns, *ns_all, *flags;
};
void launch_dosage_relayout(const uint8_t* probs, const uint8_t* miss, int64_t n_file, int bs, int rows_p,
                            const int32_t* file_idx_pad, int ref_first, uint32_t* dz, int64_t npad, cudaStream_t s);
void launch_dosage_stats(const uint32_t* dz, int64_t npad, const double* F, int dp, const int4* chunks, int nchunks,
                         int rows_p, double* part, int2* part_cnt, double* sums, double* nnz, double* n510, cudaStream_t s,
                         int ncol = 0 /* used feature columns (<= dp), 0 = all */);
void launch_s2_bt_finalize(const S2BtFinalizeArgs& a, cudaStream_t s);
// integer-unit sums [rows][4][dp] -> dosage-unit [rows][3][dp] (S1, S2, Sm) + [rows][dp] (Se)
void launch_dosage_scale(const double* sums4, int rows_p, int dp, double* sums3, double* se, cudaStream_t s);

// ---- s2_firth.cu
struct FirthArgs {
  int n_sel, C, P, dp, niter;
  double tol, maxstep;
  int64_t npad;
  const int32_t *sel_var, *sel_trait;
  const uint32_t* dz;
  const double* F;
  const double *w, *gs, *xw, *off;   // [P][Npad], xw [P][C][Npad]
  const int8_t* ym;                  // [P][Npad] 0 masked, 1 control, 2 case
  const double* xtwg;                // [bs][P][C]
  const double* mu;                  // [bs] imputed mean (after flip)
  const double* mac;                 // [bs][P]
  const int32_t* flags;              // [bs]
  double* gvec;                      // [n_sel][Npad] scratch
  int8_t* cflag;                     // [n_sel][Npad] scratch
  double *beta, *se, *lrt;
  int32_t* status;
};
void launch_s2_firth(const FirthArgs& a, cudaStream_t s);

// ---- s2_spa.cu
struct SpaArgs {
  int n_sel, C, P, dp, niter;
  double tol;
  int64_t npad;
  const int32_t *sel_var, *sel_trait;
  const uint32_t* dz;
  const double* F;
  const double *w, *gs, *xw, *phat;  // [P][Npad], xw [P][C][Npad]
  const int8_t* ym;
  const double* xtwg;                // [bs][P][C]
  const double* mu;                  // [bs]
  const double *stat, *den;          // [bs][P] score statistic and its denominator G'WG
  const int32_t* flags;
  double* gvec;                      // [n_sel][Npad] scratch
  int8_t* cflag;                     // [n_sel][Npad] scratch (active-set flags)
  double* pval;                      // [n_sel] sum of the two tail probabilities
  int32_t* status;
};
void launch_s2_spa(const SpaArgs& a, cudaStream_t s);

}  // namespace rg
