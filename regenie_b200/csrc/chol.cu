// Batched FP64 Cholesky solve of the shifted ridge systems  (A_f + lambda_r I) beta = b_f.
//
// The reference diagonalises  GGt - G_folds[f]  once per fold with SelfAdjointEigenSolver and
// applies  V (D + lambda_r I)^-1 V^T  (src/Step1_Models.cpp:484-494; level 1: :828-835).  On
// the GPU an eigendecomposition is latency-bound; the same vectors come from K*R independent
// Cholesky factorisations (n^3/3 flops each, all batched), identical to ~1e-12 relative.
//
// Storage: row-major lower triangle, ld = nC (multiple of 64); matrix m occupies
// cm + m*stride; rows nC .. nC+Ppad-1 hold the right-hand sides as extra rows, so the
// factorisation sweep leaves  y^T = (L^-1 b)^T  there (fused forward substitution).
//
// Left-looking by 64-column panels, three launches per panel:
//   chol_update : every row tile at/after the panel subtracts L[rows,0:k] L[k:k+64,0:k]^T
//                 (FP64 GEMM, the n^3/3 flops)
//   chol_diag   : one CTA per system factors the diagonal tile in registers and also produces
//                 M = L_kk^-T (the same column sweep applied to I)
//   chol_trsm   : row tiles below the panel:  L[rows, k:k+64] = P[rows, k:k+64] * M
// No CTA ever reads a tile that another CTA of the same launch overwrites: kernels of several
// "lanes" run concurrently, so launch-wide lockstep cannot be assumed.
// The stored inverses M also turn the backward substitution's triangular solves into GEMVs.
#include "gemm_dmma.cuh"
#include <stdio.h>
#include <stdlib.h>

#include "kernels.cuh"

namespace rg {

constexpr int TB = 64;  // tile / panel width

// acc[a][b] = sum_{p<k} A[r0 + ty*4+a][p] * A[rB + tx+16b][p]
__device__ __forceinline__ void gemm_tile_nt(const double* __restrict__ A, int ld, int r0, int rB, int k,
                                             double (&acc)[4][4], double (*As)[TB + 2], double (*Bs)[TB + 2]) {
  const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
  const int lrow = threadIdx.x / 4, lp = (threadIdx.x % 4) * 4;   // loader mapping: 64 rows x 16 p
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.0;
  if (k == 0) return;
  const double* arow = A + (int64_t)(r0 + lrow) * ld + lp;
  const double* brow = A + (int64_t)(rB + lrow) * ld + lp;
  double2 a0 = *reinterpret_cast<const double2*>(arow), a1 = *reinterpret_cast<const double2*>(arow + 2);
  double2 b0 = *reinterpret_cast<const double2*>(brow), b1 = *reinterpret_cast<const double2*>(brow + 2);
  for (int p0 = 0; p0 < k; p0 += 16) {
    __syncthreads();
    As[lp + 0][lrow] = a0.x; As[lp + 1][lrow] = a0.y; As[lp + 2][lrow] = a1.x; As[lp + 3][lrow] = a1.y;
    Bs[lp + 0][lrow] = b0.x; Bs[lp + 1][lrow] = b0.y; Bs[lp + 2][lrow] = b1.x; Bs[lp + 3][lrow] = b1.y;
    __syncthreads();
    if (p0 + 16 < k) {   // register prefetch of the next K-chunk
      a0 = *reinterpret_cast<const double2*>(arow + p0 + 16);
      a1 = *reinterpret_cast<const double2*>(arow + p0 + 18);
      b0 = *reinterpret_cast<const double2*>(brow + p0 + 16);
      b1 = *reinterpret_cast<const double2*>(brow + p0 + 18);
    }
#pragma unroll
    for (int p = 0; p < 16; ++p) {
      double av[4], bv[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) av[a] = As[p][ty * 4 + a];
#pragma unroll
      for (int b = 0; b < 4; ++b) bv[b] = Bs[p][tx + 16 * b];   // lane-consecutive columns: conflict-free
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = fma(av[a], bv[b], acc[a][b]);
    }
  }
}

// P[r0:r0+64, k:k+64] -= L[r0:r0+64, 0:k] * L[k:k+64, 0:k]^T   (k > 0), FP64 tensor pipe (DMMA)
// grid: (row tiles at/after the panel, 1, batch); 256 threads.
__global__ void __launch_bounds__(256)
chol_update_kernel(double* __restrict__ cm, int64_t stride, int ld, int k, int tile0) {
  __shared__ double As[TB * DM_LD];
  __shared__ double Bs[TB * DM_LD];
  double* A = cm + (int64_t)blockIdx.z * stride;
  const int r0 = (tile0 + blockIdx.x) * TB;
  DmmaAcc acc;
  gemm_tile_nt_dmma(A + (int64_t)r0 * ld, ld, true, A + (int64_t)k * ld, ld, true, k, acc, As, Bs);
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      double2* o = reinterpret_cast<double2*>(A + (int64_t)(r0 + dm_row(i)) * ld + k + dm_col(j));
      double2 v = *o;
      v.x -= acc.c[i][j][0];
      v.y -= acc.c[i][j][1];
      *o = v;
    }
}

// Diagonal tile: L_kk = chol(P_kk) and M = L_kk^-T, both in registers (thread (r, q) holds the
// columns q+4j of row r); one barrier per column, pivots broadcast through a double buffer.
// grid: (batch); 256 threads.
__global__ void __launch_bounds__(256)
chol_diag_kernel(double* __restrict__ cm, int64_t stride, int ld, int k,
                 double* __restrict__ inv, int64_t inv_stride,
                 unsigned long long* __restrict__ err_slot, long long err_base) {
  __shared__ double cb[2][TB], xb[2][TB];
  double* A = cm + (int64_t)blockIdx.x * stride;
  const int r = threadIdx.x & (TB - 1), q = threadIdx.x >> 6;
  double D[16], T[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const int cc = q + 4 * j;
    D[j] = (cc <= r) ? A[(int64_t)(k + r) * ld + k + cc] : 0.0;
    T[j] = (cc == r) ? 1.0 : 0.0;
  }
  bool bad = false;
#pragma unroll
  for (int jc = 0; jc < 16; ++jc) {
#pragma unroll
    for (int qc = 0; qc < 4; ++qc) {
      const int c = 4 * jc + qc;
      const int buf = c & 1;
      if (q == qc) { cb[buf][r] = D[jc]; xb[buf][r] = T[jc]; }
      __syncthreads();      // one barrier per column: the other buffer is only rewritten after the next barrier
      const double piv = cb[buf][c];
      if (!(piv > 0.0)) bad = true;
      const double inv_p = rsqrt(piv);
      const double lr = cb[buf][r] * inv_p;      // L[r][c]  (meaningful for r >= c)
      const double xr = xb[buf][r] * inv_p;      // X[r][c] = (L^-T)[r][c]
      if (q == qc) { D[jc] = (r >= c) ? lr : 0.0; T[jc] = xr; }
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        if (j > jc || (j == jc && q > qc)) {      // columns cc > c
          const int cc = q + 4 * j;
          const double lcc = cb[buf][cc] * inv_p;
          if (r >= cc) D[j] = fma(-lr, lcc, D[j]);
          T[j] = fma(-xr, lcc, T[j]);
        }
      }
    }
  }
  if (bad && threadIdx.x == 0) atomicMin(err_slot, (unsigned long long)(err_base + blockIdx.x + 1));
  double* Mo = inv + (int64_t)blockIdx.x * inv_stride + (int64_t)(k / TB) * TB * TB;
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const int cc = q + 4 * j;
    if (cc <= r) A[(int64_t)(k + r) * ld + k + cc] = D[j];
    Mo[r * TB + cc] = (cc >= r) ? T[j] : 0.0;
  }
}

// L[rows, k:k+64] = P[rows, k:k+64] * M,  M = L_kk^-T (upper triangular).
// grid: (row tiles strictly below the panel, 1, batch); 256 threads, 4x4 outputs each.
__global__ void __launch_bounds__(256)
chol_trsm_kernel(double* __restrict__ cm, int64_t stride, int ld, int k, int tile0,
                 const double* __restrict__ inv, int64_t inv_stride) {
  extern __shared__ double trsm_sm[];
  double (*Ts)[TB + 1] = reinterpret_cast<double (*)[TB + 1]>(trsm_sm);
  double (*Ms)[TB + 2] = reinterpret_cast<double (*)[TB + 2]>(trsm_sm + TB * (TB + 1));
  double* A = cm + (int64_t)blockIdx.z * stride;
  const double* M = inv + (int64_t)blockIdx.z * inv_stride + (int64_t)(k / TB) * TB * TB;
  const int r0 = (tile0 + 1 + blockIdx.x) * TB;
  for (int e = threadIdx.x; e < TB * TB; e += 256) {
    const int rr = e / TB, cc = e % TB;
    Ts[rr][cc] = A[(int64_t)(r0 + rr) * ld + k + cc];
    Ms[rr][cc] = M[e];
  }
  __syncthreads();
  const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
  double acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.0;
#pragma unroll 8
  for (int p = 0; p < TB; ++p) {
    double av[4], bv[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) av[a] = Ts[ty * 4 + a][p];
#pragma unroll
    for (int b = 0; b < 4; ++b) bv[b] = Ms[p][tx + 16 * b];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) acc[a][b] = fma(av[a], bv[b], acc[a][b]);
  }
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    double* o = A + (int64_t)(r0 + ty * 4 + a) * ld + k + tx;
#pragma unroll
    for (int b = 0; b < 4; ++b) o[16 * b] = acc[a][b];
  }
}

// Backward substitution  L^T beta = y  for all right-hand sides, one 1024-thread CTA per matrix,
// sweeping 64-column blocks from the last to the first.  y lives in the RHS rows (row nC+p,
// contiguous over columns); beta overwrites it.  The diagonal solves are GEMVs with the stored
// M = L_kk^-T; the sweep  y[:, j] -= L[k+r][j] beta[r]  streams the 64 panel rows once.
constexpr int BS_THREADS = 256;
constexpr int BS_JT = 4;        // columns per thread in the sweep (amortises the broadcast beta loads)
constexpr int BS_PC = 10;   // right-hand sides per register pass

__global__ void __launch_bounds__(BS_THREADS)
chol_backsolve_kernel(double* __restrict__ cm, int64_t stride, int ld, int nC, int P,
                      const double* __restrict__ inv, int64_t inv_stride) {
  extern __shared__ double back_sm[];
  double* Ms = back_sm;                    // [TB][TB+1]
  double* ys = Ms + TB * (TB + 1);         // [TB][P]  y block
  double* bs = ys + TB * P;                // [TB][P]  beta block
  double* A = cm + (int64_t)blockIdx.x * stride;
  const double* Minv = inv + (int64_t)blockIdx.x * inv_stride;
  for (int kb = nC / TB - 1; kb >= 0; --kb) {
    const int k = kb * TB;
    __syncthreads();
    for (int e = threadIdx.x; e < TB * TB; e += BS_THREADS)
      Ms[(e / TB) * (TB + 1) + (e % TB)] = Minv[(int64_t)kb * TB * TB + e];
    for (int e = threadIdx.x; e < TB * P; e += BS_THREADS) {
      const int r = e % TB, p = e / TB;
      ys[r * P + p] = A[(int64_t)(nC + p) * ld + k + r];
    }
    __syncthreads();
    // beta = L_kk^-T y = M y   (M upper triangular: M[r][c], c >= r)
    for (int e = threadIdx.x; e < TB * P; e += BS_THREADS) {
      const int r = e % TB, p = e / TB;
      double s = 0.0;
      for (int c = r; c < TB; ++c) s = fma(Ms[r * (TB + 1) + c], ys[c * P + p], s);
      bs[r * P + p] = s;
      A[(int64_t)(nC + p) * ld + k + r] = s;
    }
    __syncthreads();
    // y[p][j] -= sum_r L[k+r][j] * beta[r][p]   for all j < k   (coalesced over j; each thread owns BS_JT columns)
    for (int j0 = threadIdx.x; j0 < k; j0 += BS_THREADS * BS_JT) {
      for (int p0 = 0; p0 < P; p0 += BS_PC) {
        const int np = min(BS_PC, P - p0);
        double acc[BS_JT][BS_PC];
#pragma unroll
        for (int m = 0; m < BS_JT; ++m)
#pragma unroll
          for (int qq = 0; qq < BS_PC; ++qq) acc[m][qq] = 0.0;
#pragma unroll 1
        for (int rb = 0; rb < TB; rb += 8) {
          double l[BS_JT][8];
#pragma unroll
          for (int m = 0; m < BS_JT; ++m) {
            const int j = j0 + m * BS_THREADS;
#pragma unroll
            for (int u = 0; u < 8; ++u) l[m][u] = (j < k) ? A[(int64_t)(k + rb + u) * ld + j] : 0.0;
          }
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const double* bb = bs + (rb + u) * P + p0;
#pragma unroll
            for (int qq = 0; qq < BS_PC; ++qq) {
              const double bv = (qq < np) ? bb[qq] : 0.0;
#pragma unroll
              for (int m = 0; m < BS_JT; ++m) acc[m][qq] = fma(l[m][u], bv, acc[m][qq]);
            }
          }
        }
#pragma unroll
        for (int m = 0; m < BS_JT; ++m) {
          const int j = j0 + m * BS_THREADS;
          if (j < k)
#pragma unroll
            for (int qq = 0; qq < BS_PC; ++qq)
              if (qq < np) A[(int64_t)(nC + p0 + qq) * ld + j] -= acc[m][qq];
        }
      }
    }
  }
}

void launch_chol_factor(double* cm, int64_t stride, int nC, int n_aug, int batch, double* inv,
                        unsigned long long* err_slot, long long err_base, cudaStream_t s) {
  const int ntiles = n_aug / TB;
  const int64_t inv_stride = (int64_t)(nC / TB) * TB * TB;
  const size_t trsm_smem = ((size_t)TB * (TB + 1) + (size_t)TB * (TB + 2)) * sizeof(double);
  static bool attr_set = false;
  if (!attr_set) {
    RG_CUDA(cudaFuncSetAttribute(chol_trsm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)trsm_smem));
    attr_set = true;
  }
  // profiling aid (RG_B200_CHOL_TIMING=1): CUDA-event time of the three kernels of every panel step
  static const bool timing = getenv("RG_B200_CHOL_TIMING") != nullptr;
  static double t_acc[3] = {0, 0, 0};
  static long t_calls = 0;
  cudaEvent_t ev[4];
  if (timing) for (auto& e : ev) cudaEventCreate(&e);
  auto tick = [&](int i) { if (timing) cudaEventRecord(ev[i], s); };
  auto tock = [&]() {
    if (!timing) return;
    cudaEventSynchronize(ev[3]);
    for (int i = 0; i < 3; ++i) { float ms = 0; cudaEventElapsedTime(&ms, ev[i], ev[i + 1]); t_acc[i] += ms; }
  };
  for (int kb = 0; kb < nC / TB; ++kb) {
    const int k = kb * TB;
    dim3 g1(ntiles - kb, 1, batch);
    tick(0);
    if (k > 0) chol_update_kernel<<<g1, 256, 0, s>>>(cm, stride, nC, k, kb);
    tick(1);
    chol_diag_kernel<<<batch, 256, 0, s>>>(cm, stride, nC, k, inv, inv_stride, err_slot, err_base);
    tick(2);
    dim3 g2(ntiles - kb - 1, 1, batch);
    if (ntiles - kb - 1 > 0) chol_trsm_kernel<<<g2, 256, trsm_smem, s>>>(cm, stride, nC, k, kb, inv, inv_stride);
    tick(3);
    tock();
  }
  if (timing) {
    for (auto& e : ev) cudaEventDestroy(e);
    if (++t_calls % 50 == 0)
      fprintf(stderr, "[chol timing] per factorisation (ms): update %.3f diag %.3f trsm %.3f\n", t_acc[0] / t_calls,
              t_acc[1] / t_calls, t_acc[2] / t_calls);
  }
}

void launch_chol_backsolve(double* cm, int64_t stride, int nC, int P, int batch, const double* inv,
                           cudaStream_t s) {
  const size_t smem = ((size_t)TB * (TB + 1) + (size_t)2 * TB * P) * sizeof(double);
  static size_t smem_set = 0;
  if (smem > smem_set) {
    RG_CUDA(cudaFuncSetAttribute(chol_backsolve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    smem_set = smem;
  }
  const int64_t inv_stride = (int64_t)(nC / TB) * TB * TB;
  chol_backsolve_kernel<<<batch, BS_THREADS, smem, s>>>(cm, stride, nC, nC, P, inv, inv_stride);
}

int chol_num_launches(int nC) { return 3 * (nC / TB) - 1; }
size_t chol_inv_elems(int nC, int batch) { return (size_t)batch * (nC / TB) * TB * TB; }

}  // namespace rg

namespace rg {

// Row-wise backward substitution for MANY right-hand sides stored as rows (LOOCV):
//   rows hold t_i^T = (L^-1 w_i)^T;  on exit they hold z_i^T = t_i^T L^-1 = (L^-T t_i)^T = (H w_i)^T.
// grid: (row tiles, 1, batch); one CTA owns 64 rows and sweeps the column blocks from last to first:
//   z_kb = t_kb * L_kk^-1 (= t_kb * M^T),   t[:, 0:k] -= z_kb * L[k:k+64, 0:k].
__global__ void __launch_bounds__(256)
chol_rows_backsolve_kernel(double* __restrict__ cm, int64_t stride, int ld, int nC, int row0,
                           const double* __restrict__ inv, int64_t inv_stride) {
  extern __shared__ double rb_sm[];
  double (*Zs)[TB + 1] = reinterpret_cast<double (*)[TB + 1]>(rb_sm);                     // z_kb / t_kb tile
  double (*Ls)[TB + 2] = reinterpret_cast<double (*)[TB + 2]>(rb_sm + TB * (TB + 1));     // M or an L tile
  double* A = cm + (int64_t)blockIdx.z * stride;
  const double* Minv = inv + (int64_t)blockIdx.z * inv_stride;
  const int r0 = row0 + blockIdx.x * TB;
  const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
  for (int kb = nC / TB - 1; kb >= 0; --kb) {
    const int k = kb * TB;
    __syncthreads();
    for (int e = threadIdx.x; e < TB * TB; e += 256) {
      const int rr = e / TB, cc = e % TB;
      Zs[rr][cc] = A[(int64_t)(r0 + rr) * ld + k + cc];
      Ls[rr][cc] = Minv[(int64_t)kb * TB * TB + e];          // M[c][p] at Ls[c][p]
    }
    __syncthreads();
    // z[row][c] = sum_p t[row][p] * M[c][p]
    double z[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) z[a][b] = 0.0;
#pragma unroll 8
    for (int p = 0; p < TB; ++p) {
      double av[4], bv[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) av[a] = Zs[ty * 4 + a][p];
#pragma unroll
      for (int b = 0; b < 4; ++b) bv[b] = Ls[tx + 16 * b][p];
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) z[a][b] = fma(av[a], bv[b], z[a][b]);
    }
    __syncthreads();
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        Zs[ty * 4 + a][tx + 16 * b] = z[a][b];
        A[(int64_t)(r0 + ty * 4 + a) * ld + k + tx + 16 * b] = z[a][b];
      }
    // t[:, j0:j0+64] -= z_kb * L[k:k+64, j0:j0+64]
    for (int j0 = 0; j0 < k; j0 += TB) {
      __syncthreads();
      for (int e = threadIdx.x; e < TB * TB; e += 256) {
        const int rr = e / TB, cc = e % TB;
        Ls[rr][cc] = A[(int64_t)(k + rr) * ld + j0 + cc];    // L[k+p][j0+c] at Ls[p][c]
      }
      __syncthreads();
      double u[4][4];
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) u[a][b] = 0.0;
#pragma unroll 8
      for (int p = 0; p < TB; ++p) {
        double av[4], bv[4];
#pragma unroll
        for (int a = 0; a < 4; ++a) av[a] = Zs[ty * 4 + a][p];
#pragma unroll
        for (int b = 0; b < 4; ++b) bv[b] = Ls[p][tx + 16 * b];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int b = 0; b < 4; ++b) u[a][b] = fma(av[a], bv[b], u[a][b]);
      }
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) A[(int64_t)(r0 + ty * 4 + a) * ld + j0 + tx + 16 * b] -= u[a][b];
    }
  }
}

void launch_chol_rows_backsolve(double* cm, int64_t stride, int nC, int row0, int nrows, int batch,
                                const double* inv, cudaStream_t s) {
  const size_t smem = ((size_t)TB * (TB + 1) + (size_t)TB * (TB + 2)) * sizeof(double);
  static bool attr_set = false;
  if (!attr_set) {
    RG_CUDA(cudaFuncSetAttribute(chol_rows_backsolve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  const int64_t inv_stride = (int64_t)(nC / TB) * TB * TB;
  dim3 grid(nrows / TB, 1, batch);
  chol_rows_backsolve_kernel<<<grid, 256, smem, s>>>(cm, stride, nC, nC, row0, inv, inv_stride);
}

}  // namespace rg
