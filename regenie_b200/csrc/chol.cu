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
// Left-looking by 64-column panels: trailing data is read, never rewritten -> n^3/6 * 8 / 64
// bytes of traffic per matrix instead of n^3/3 * 8 / 64 * 2 for right-looking.
#include "kernels.cuh"

namespace rg {

constexpr int TB = 64;  // tile / panel width

// P[r0:r0+64, k:k+64] -= L[r0:r0+64, 0:k] * L[k:k+64, 0:k]^T      (k > 0)
// grid: (row tiles at/after the panel, 1, batch); 256 threads, 4x4 register tile each.
__global__ void __launch_bounds__(256)
chol_update_kernel(double* __restrict__ cm, int64_t stride, int ld, int k, int tile0) {
  __shared__ double As[16][TB + 2];
  __shared__ double Bs[16][TB + 2];
  double* A = cm + (int64_t)blockIdx.z * stride;
  const int r0 = (tile0 + blockIdx.x) * TB;
  const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
  const int lrow = threadIdx.x / 4, lp = (threadIdx.x % 4) * 4;   // loader mapping: 64 rows x 16 p
  double acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.0;

  const double* arow = A + (int64_t)(r0 + lrow) * ld + lp;
  const double* brow = A + (int64_t)(k + lrow) * ld + lp;
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
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    double* o = A + (int64_t)(r0 + ty * 4 + a) * ld + k + tx;
#pragma unroll
    for (int b = 0; b < 4; ++b) o[16 * b] -= acc[a][b];
  }
}

// Panel step: every CTA factors the (already updated) 64x64 diagonal block redundantly in
// shared memory and applies the same column sweep to its own 64-row tile:
//   L_kk = chol(P_kk);   L[r0:r0+64, k:k+64] = P[r0.., k..] * L_kk^-T.
// grid: (row tiles at/after the panel, 1, batch); tile 0 is the diagonal block itself.
__global__ void __launch_bounds__(256)
chol_panel_kernel(double* __restrict__ cm, int64_t stride, int ld, int k, int tile0,
                  unsigned long long* __restrict__ err_slot, long long err_base) {
  extern __shared__ double panel_sm[];
  double (*D)[TB + 1] = reinterpret_cast<double (*)[TB + 1]>(panel_sm);
  double (*T)[TB + 1] = reinterpret_cast<double (*)[TB + 1]>(panel_sm + TB * (TB + 1));
  double* A = cm + (int64_t)blockIdx.z * stride;
  const bool is_diag = (blockIdx.x == 0);
  const int r0 = (tile0 + blockIdx.x) * TB;
  for (int e = threadIdx.x; e < TB * TB; e += 256) {
    const int r = e / TB, c = e % TB;
    D[r][c] = (c <= r) ? A[(int64_t)(k + r) * ld + k + c] : 0.0;
    if (!is_diag) T[r][c] = A[(int64_t)(r0 + r) * ld + k + c];
  }
  __syncthreads();
  for (int c = 0; c < TB; ++c) {
    const double piv = D[c][c];
    if (!(piv > 0.0) && threadIdx.x == 0 && is_diag)
      atomicMin(err_slot, (unsigned long long)(err_base + blockIdx.z + 1));
    const double inv = 1.0 / sqrt(piv);
    __syncthreads();
    // scale column c
    if (threadIdx.x < TB) {
      const int r = threadIdx.x;
      if (r >= c) D[r][c] *= inv;           // r == c: piv/sqrt(piv) = sqrt(piv)
    } else if (threadIdx.x < 2 * TB && !is_diag) {
      T[threadIdx.x - TB][c] *= inv;
    }
    __syncthreads();
    // rank-1 update of the remaining columns: thread = (row, column phase), no divisions
    {
      const int r = threadIdx.x & (TB - 1);
      const double dr = D[r][c];
      const double tr = is_diag ? 0.0 : T[r][c];
      for (int cc = c + 1 + (threadIdx.x >> 6); cc < TB; cc += 4) {
        const double l = D[cc][c];
        if (r >= cc) D[r][cc] -= dr * l;
        if (!is_diag) T[r][cc] -= tr * l;
      }
    }
    // (the next iteration's first __syncthreads orders these writes before the column scale)
    __syncthreads();
  }
  for (int e = threadIdx.x; e < TB * TB; e += 256) {
    const int r = e / TB, c = e % TB;
    if (is_diag) {
      if (c <= r) A[(int64_t)(k + r) * ld + k + c] = D[r][c];
    } else {
      A[(int64_t)(r0 + r) * ld + k + c] = T[r][c];
    }
  }
}

// Backward substitution  L^T beta = y  for all right-hand sides, one CTA per matrix,
// sweeping 64-column blocks from the last to the first ("right-looking" on y).
// y lives in the RHS rows (row nC+p, contiguous over columns); beta overwrites it.
__global__ void __launch_bounds__(256)
chol_backsolve_kernel(double* __restrict__ cm, int64_t stride, int ld, int nC, int P) {
  extern __shared__ double back_sm[];
  double (*Lkk)[TB + 1] = reinterpret_cast<double (*)[TB + 1]>(back_sm);
  double* bsm = back_sm + TB * (TB + 1);   // y / beta of the current block: [TB][P]
  double* A = cm + (int64_t)blockIdx.x * stride;
  for (int kb = nC / TB - 1; kb >= 0; --kb) {
    const int k = kb * TB;
    __syncthreads();
    for (int e = threadIdx.x; e < TB * TB; e += 256) {
      const int r = e / TB, c = e % TB;
      Lkk[r][c] = (c <= r) ? A[(int64_t)(k + r) * ld + k + c] : 0.0;
    }
    for (int e = threadIdx.x; e < TB * P; e += 256) {
      const int r = e % TB, p = e / TB;
      bsm[r * P + p] = A[(int64_t)(nC + p) * ld + k + r];
    }
    // column-oriented solve of L_kk^T x = y_k: one row per step, the rest updated in parallel
    for (int r = TB - 1; r >= 0; --r) {
      __syncthreads();
      if (threadIdx.x < P) bsm[r * P + threadIdx.x] /= Lkk[r][r];
      __syncthreads();
      for (int e = threadIdx.x; e < r * P; e += 256) {
        const int rr = e / P, p = e - rr * P;
        bsm[rr * P + p] -= Lkk[r][rr] * bsm[r * P + p];
      }
    }
    __syncthreads();
    for (int e = threadIdx.x; e < TB * P; e += 256) {
      const int r = e % TB, p = e / TB;
      A[(int64_t)(nC + p) * ld + k + r] = bsm[r * P + p];
    }
    // y[p][j] -= sum_r L[k+r][j] * beta[r][p]   for all j < k   (coalesced over j)
    for (int j = threadIdx.x; j < k; j += 256) {
      for (int p0 = 0; p0 < P; p0 += 10) {
        double acc[10];
#pragma unroll
        for (int q = 0; q < 10; ++q) acc[q] = 0.0;
        const int np = min(10, P - p0);
#pragma unroll 8
        for (int r = 0; r < TB; ++r) {
          const double l = A[(int64_t)(k + r) * ld + j];
          const double* bb = bsm + r * P + p0;
          if (np == 10) {
#pragma unroll
            for (int q = 0; q < 10; ++q) acc[q] = fma(l, bb[q], acc[q]);
          } else {
            for (int q = 0; q < np; ++q) acc[q] = fma(l, bb[q], acc[q]);
          }
        }
        for (int q = 0; q < np; ++q) A[(int64_t)(nC + p0 + q) * ld + j] -= acc[q];
      }
    }
  }
}

void launch_chol_factor(double* cm, int64_t stride, int nC, int n_aug, int batch,
                        unsigned long long* err_slot, long long err_base, cudaStream_t s) {
  const int ntiles = n_aug / TB;
  const size_t panel_smem = (size_t)2 * TB * (TB + 1) * sizeof(double);
  static bool attr_set = false;
  if (!attr_set) {
    RG_CUDA(cudaFuncSetAttribute(chol_panel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)panel_smem));
    attr_set = true;
  }
  for (int kb = 0; kb < nC / TB; ++kb) {
    const int k = kb * TB;
    dim3 grid(ntiles - kb, 1, batch);
    if (k > 0) chol_update_kernel<<<grid, 256, 0, s>>>(cm, stride, nC, k, kb);
    chol_panel_kernel<<<grid, 256, panel_smem, s>>>(cm, stride, nC, k, kb, err_slot, err_base);
  }
}

void launch_chol_backsolve(double* cm, int64_t stride, int nC, int P, int batch, cudaStream_t s) {
  const size_t smem = ((size_t)TB * (TB + 1) + (size_t)TB * P) * sizeof(double);
  RG_CUDA(cudaFuncSetAttribute(chol_backsolve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  chol_backsolve_kernel<<<batch, 256, smem, s>>>(cm, stride, nC, nC, P);
}

int chol_num_launches(int nC) { return 2 * (nC / TB) - 1 + 1; }

}  // namespace rg
