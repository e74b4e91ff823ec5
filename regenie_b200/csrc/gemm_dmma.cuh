// 64x64 FP64 "NT" tile  C = A_rows * B_rows^T  on the FP64 tensor pipe (mma.sync m8n8k4 DMMA).
// Both operands are row tiles with the contraction index contiguous (rows of L, or columns of W):
//   acc(m, n) = sum_p Arow[m][p] * Brow[n][p]
// 256 threads = 8 warps; warp (wm, wn) owns a 16 x 32 sub-tile = 2 x 4 DMMA tiles; per thread
// 16 accumulators.  Fragment loads are bank-conflict free with a row stride of 20 doubles.
#pragma once
#include <stdint.h>

namespace rg {

constexpr int DM_KC = 16;          // contraction chunk staged per barrier pair
constexpr int DM_LD = DM_KC + 4;   // smem row stride (doubles): stride mod 16 == 4 -> conflict-free fragments

struct DmmaAcc {
  double c[2][4][2];               // [m tile][n tile][2 consecutive columns]
};

__device__ __forceinline__ void dmma_8x8x4(double (&d)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
               : "+d"(d[0]), "+d"(d[1]) : "d"(a), "d"(b));
}

// ap/bp: THIS thread's loader pointers (row threadIdx.x >> 2 of the tile, element (threadIdx.x & 3) * 4 of the row), so the
// rows of an operand need not be equally spaced; va/vb: per-loader-row validity (invalid rows are read as zero).
// klen must be a multiple of 16.  As/Bs: [64][DM_LD] doubles each.
__device__ __forceinline__ void gemm_tile_nt_dmma_ptr(const double* __restrict__ ap, bool va, const double* __restrict__ bp,
                                                      bool vb, int klen, DmmaAcc& acc, double* As, double* Bs) {
  const int lrow = threadIdx.x >> 2, lp = (threadIdx.x & 3) * 4;     // loader: 64 rows x 16 p
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wm = warp >> 1, wn = warp & 1;
  const int fr = lane >> 2, fk = lane & 3;                           // fragment row / k within the 8x4 tile
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc.c[i][j][0] = acc.c[i][j][1] = 0.0;
  if (klen <= 0) return;
  const double2 z2 = make_double2(0.0, 0.0);
  double2 a0 = va ? *reinterpret_cast<const double2*>(ap) : z2, a1 = va ? *reinterpret_cast<const double2*>(ap + 2) : z2;
  double2 b0 = vb ? *reinterpret_cast<const double2*>(bp) : z2, b1 = vb ? *reinterpret_cast<const double2*>(bp + 2) : z2;
  for (int p0 = 0; p0 < klen; p0 += DM_KC) {
    __syncthreads();
    *reinterpret_cast<double2*>(As + lrow * DM_LD + lp) = a0;
    *reinterpret_cast<double2*>(As + lrow * DM_LD + lp + 2) = a1;
    *reinterpret_cast<double2*>(Bs + lrow * DM_LD + lp) = b0;
    *reinterpret_cast<double2*>(Bs + lrow * DM_LD + lp + 2) = b1;
    __syncthreads();
    if (p0 + DM_KC < klen) {     // register prefetch of the next chunk
      if (va) { a0 = *reinterpret_cast<const double2*>(ap + p0 + DM_KC); a1 = *reinterpret_cast<const double2*>(ap + p0 + DM_KC + 2); }
      if (vb) { b0 = *reinterpret_cast<const double2*>(bp + p0 + DM_KC); b1 = *reinterpret_cast<const double2*>(bp + p0 + DM_KC + 2); }
    }
#pragma unroll
    for (int kk = 0; kk < DM_KC; kk += 4) {
      double af[2], bf[4];
#pragma unroll
      for (int i = 0; i < 2; ++i) af[i] = As[(wm * 16 + i * 8 + fr) * DM_LD + kk + fk];
#pragma unroll
      for (int j = 0; j < 4; ++j) bf[j] = Bs[(wn * 32 + j * 8 + fr) * DM_LD + kk + fk];
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) dmma_8x8x4(acc.c[i][j], af[i], bf[j]);
    }
  }
}

// arow/brow: pointers to element [tile row 0][0]; lda/ldb: row strides (doubles).
__device__ __forceinline__ void gemm_tile_nt_dmma(const double* __restrict__ arow, int64_t lda, bool va,
                                                  const double* __restrict__ brow, int64_t ldb, bool vb, int klen,
                                                  DmmaAcc& acc, double* As, double* Bs) {
  const int lrow = threadIdx.x >> 2, lp = (threadIdx.x & 3) * 4;
  gemm_tile_nt_dmma_ptr(arow + (int64_t)lrow * lda + lp, va, brow + (int64_t)lrow * ldb + lp, vb, klen, acc, As, Bs);
}

// element (i, j, e) of the accumulator sits at tile row dm_row(i), tile column dm_col(j) + e
__device__ __forceinline__ int dm_row(int i) { return ((threadIdx.x >> 5) >> 1) * 16 + i * 8 + ((threadIdx.x & 31) >> 2); }
__device__ __forceinline__ int dm_col(int j) { return ((threadIdx.x >> 5) & 1) * 32 + j * 8 + ((threadIdx.x & 31) & 3) * 2; }

}  // namespace rg
