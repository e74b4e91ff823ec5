// Out-of-fold level-0 predictions on the tensor cores, exactly.
//
//   pred[t, q] = sum_i gamma[i,q] g0(i,t) + sum_i (gamma mu)[i,q] miss(i,t) - x_t . cvec_q
// (reference: `beta.transpose() * Gmat.block(...)`, src/Step1_Models.cpp:503 - 2 P R bs N flops/block).
// The genotype operand is the same e4m3 plane pair Z = [G0; Miss] the Gram kernel consumes (exact
// small integers).  The real-valued coefficients are split into 9 balanced radix-30 digits
//   gamma[i,q] = (s_q / 15) * sum_l d_l[i,q] 30^-l,   d_l in {-15..15}  (exact in e4m3),
// so every tensor-core product is an integer <= 30 and every K = 2*rows_p sum is < 2^24: the FP32
// TMEM accumulators hold EXACT integers and the FP64 epilogue reassembles the prediction to
// ~5e-14 * s_q per coefficient (44 bits), far inside the 1e-5 parity budget.
//
// Orientation: samples are the MMA M dimension (128 samples = 128 TMEM lanes), the 9 x 56 digit
// rows are N (2 x 256 columns = all 512 TMEM columns), the SNP/plane index is K.
//   A = Z^T tile: 128 samples x K, "MN-major" (samples contiguous) straight from the Z rows
//   B = digit rows: 512 x K, K-major
// so each epilogue thread owns one sample and all of its (limb, q) partial sums: no cross-lane
// transpose, the limbs are folded in registers.
#include <stdlib.h>

#include "kernels.cuh"

namespace rg {

namespace {

constexpr int PT_BM = 128;            // samples per CTA
constexpr int PT_BN = 512;            // digit rows (9 limbs x 56 outputs, zero padded)
constexpr int PT_BK = 128;            // Z rows per stage
constexpr int PT_STAGES = 2;
constexpr int PT_A_BYTES = PT_BK * PT_BM;     // 16 KiB: 128 k-rows x 128 samples
constexpr int PT_B_BYTES = PT_BN * PT_BK;     // 64 KiB: 512 digit rows x 128 k bytes
constexpr int PT_STAGE_BYTES = PT_A_BYTES + PT_B_BYTES;
constexpr int PT_THREADS = 320;          // TMA warp, MMA warp, 8 epilogue warps (2 per TMEM lane quarter)
constexpr int PT_QH = kLimbQ / 2;        // outputs per epilogue thread
constexpr uint32_t PT_SPIN_LIMIT = 1u << 28;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0, spins = 0;
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (done) break;
    if (++spins > PT_SPIN_LIMIT) __trap();
  }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(tm), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_prefetch_2d(const CUtensorMap* tm, int c0, int c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(tm), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tcgen05_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// shared-memory matrix descriptor, 128B swizzle, 8-row groups 1024 B apart (both majors use it:
// K-major rows = MN index, MN-major rows = K index with the 128 contiguous bytes along MN)
__device__ __forceinline__ uint64_t make_desc(uint32_t addr) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// kind::f8f6f4, E4M3 x E4M3 -> F32, A MN-major (bit 15), B K-major, M = 128, N = 256
constexpr uint32_t kPtIdesc = (1u << 4) | (1u << 15) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

__device__ __forceinline__ void mma_f8(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(da), "l"(db), "r"(kPtIdesc), "r"(accumulate) : "memory");
}

__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// e4m3 encodings of the integers 0..15
__device__ __constant__ uint8_t kE4m3Int[16] = {0x00, 0x38, 0x40, 0x44, 0x48, 0x4A, 0x4C, 0x4E,
                                                0x50, 0x51, 0x52, 0x53, 0x54, 0x55, 0x56, 0x57};

}  // namespace

// ---------------------------------------------------------------------------------------------
// Column scales  s[f][q] = max_i max(|gamma|, |gamma mu|)  and the radix-30 digit rows
//   dig[f][g][l*56 + qq][k],  k = i (G0 plane) or rows_p + i (Miss plane),  q = g*56 + qq.
// grid: (Qp, K folds), block 256: one CTA per (column, fold).
__global__ void __launch_bounds__(256)
l0_gamma_limbs_kernel(const double* __restrict__ gam, const double* __restrict__ gmu, int Qp, int Q, int bs,
                      int rows_p, double* __restrict__ scale, uint8_t* __restrict__ dig, int ngroups) {
  __shared__ double red[256];
  const int q = blockIdx.x, f = blockIdx.y;
  if (q >= Q) return;
  const double* gcol = gam + (int64_t)f * rows_p * Qp + q;
  const double* mcol = gmu + (int64_t)f * rows_p * Qp + q;
  double mx = 0.0;
  for (int i = threadIdx.x; i < bs; i += 256)
    mx = fmax(mx, fmax(fabs(gcol[(int64_t)i * Qp]), fabs(mcol[(int64_t)i * Qp])));
  red[threadIdx.x] = mx;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] = fmax(red[threadIdx.x], red[threadIdx.x + o]);
    __syncthreads();
  }
  const double s = red[0] > 0.0 ? red[0] : 1.0;
  if (threadIdx.x == 0) scale[(int64_t)f * Qp + q] = s;
  const int g = q / kLimbQ, qq = q % kLimbQ;
  const int K2 = 2 * rows_p;
  uint8_t* base = dig + ((int64_t)f * ngroups + g) * (int64_t)PT_BN * K2;
  for (int k = threadIdx.x; k < K2; k += 256) {
    const int plane = k >= rows_p, i = plane ? k - rows_p : k;
    double v = 0.0;
    if (i < bs) v = (plane ? mcol[(int64_t)i * Qp] : gcol[(int64_t)i * Qp]) / s * 15.0;
#pragma unroll
    for (int l = 0; l < kLimbs; ++l) {
      const double d = rint(v);
      const int di = (int)d;
      base[(int64_t)(l * kLimbQ + qq) * K2 + k] = (uint8_t)(kE4m3Int[di < 0 ? -di : di] | (di < 0 ? 0x80 : 0));
      v = (v - d) * 30.0;
    }
  }
}

// grid: (Npad/128 sample tiles, q groups); 192 threads.
__global__ void __launch_bounds__(PT_THREADS, 1)
l0_predict_tcgen05_kernel(const __grid_constant__ CUtensorMap tmZ, const __grid_constant__ CUtensorMap tmD,
                          PredictTcArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gen_base = smem_raw + (base - raw);
  const uint32_t sA = base;
  const uint32_t sB = base + PT_STAGES * PT_A_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(gen_base + PT_STAGES * PT_STAGE_BYTES);
  const uint32_t full_bar = smem_u32(bars);
  const uint32_t empty_bar = smem_u32(bars + PT_STAGES);
  const uint32_t tmem_full_bar = smem_u32(bars + 2 * PT_STAGES);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * PT_STAGES + 1);
  double* s_scale = reinterpret_cast<double*>(bars + 2 * PT_STAGES + 2);   // [kLimbQ]
  double* s_cvec = s_scale + kLimbQ;                                        // [kLimbQ][C]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = blockIdx.x, g = blockIdx.y;
  const int f = a.tile_fold[tile];
  const int nkb = (2 * a.rows_p) / PT_BK;
  const int q0 = g * kLimbQ;
  const int nq = min(kLimbQ, a.Q - q0);

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < PT_STAGES; ++s) { mbar_init(full_bar + 8 * s, 1); mbar_init(empty_bar + 8 * s, 1); }
    mbar_init(tmem_full_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // epilogue constants
  for (int e = threadIdx.x; e < kLimbQ; e += PT_THREADS)
    s_scale[e] = (e < nq) ? a.scale[(int64_t)f * a.Qp + q0 + e] / 15.0 : 0.0;
  for (int e = threadIdx.x; e < kLimbQ * a.C; e += PT_THREADS) {
    const int qq = e / a.C, c = e % a.C;
    s_cvec[e] = (qq < nq) ? a.cvec[((int64_t)f * a.Qp + q0 + qq) * a.C + c] : 0.0;
  }
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const long long t_start = clock64();

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer: A = 128 Z rows x 128 samples; B = 512 digit rows x 128 k bytes =====
      const int drow0 = (f * a.ngroups + g) * PT_BN;
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % PT_STAGES;
        const uint32_t ph = (kb / PT_STAGES) & 1;
        mbar_wait(empty_bar + 8 * s, ph ^ 1);
        mbar_expect_tx(full_bar + 8 * s, PT_STAGE_BYTES);
        tma_load_2d(sA + s * PT_A_BYTES, &tmZ, full_bar + 8 * s, tile * PT_BM, kb * PT_BK);
#pragma unroll
        for (int j = 0; j < 4; ++j)
          tma_load_2d(sB + s * PT_B_BYTES + j * 16384, &tmD, full_bar + 8 * s, kb * PT_BK, drow0 + j * 128);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===== MMA issuer =====
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % PT_STAGES;
        const uint32_t ph = (kb / PT_STAGES) & 1;
        mbar_wait(full_bar + 8 * s, ph);
        fence_after();
        const uint64_t da = make_desc(sA + s * PT_A_BYTES);
        const uint64_t db0 = make_desc(sB + s * PT_B_BYTES);
        const uint64_t db1 = make_desc(sB + s * PT_B_BYTES + 256 * 128);
#pragma unroll
        for (int k = 0; k < PT_BK / 32; ++k) {
          // A (MN-major): 32 k-rows = 4 groups of 1024 B -> +4096 B per step (256 in 16-byte units)
          // B (K-major) : +32 bytes inside the 128-byte swizzle atom (2 in 16-byte units)
          const uint32_t acc = (kb | k) ? 1u : 0u;
          mma_f8(tmem_base, da + (uint64_t)(256 * k), db0 + (uint64_t)(2 * k), acc);
          mma_f8(tmem_base + 256u, da + (uint64_t)(256 * k), db1 + (uint64_t)(2 * k), acc);
        }
        tcgen05_commit(empty_bar + 8 * s);
      }
      tcgen05_commit(tmem_full_bar);
    }
  } else {
    // ===== epilogue: thread = (sample, half of the outputs).  The limb sums are exact integers < 2^22:
    // convert with the 1.5*2^23 magic add, fold limbs 0-3 / 4-7 / 8 in int32 Horner form, then 3 FP64 FMAs.
    const int qw = warp & 3;                      // TMEM lane quarter this warp may access
    const int half = (warp - 2) >> 2;             // which 28 of the 56 outputs
    const int t = tile * PT_BM + qw * 32 + lane;
    mbar_wait(tmem_full_bar, 0);
    fence_after();
    const long long t_mma = clock64();
    int ihi[PT_QH], imid[PT_QH], ilo[PT_QH];
#pragma unroll
    for (int j = 0; j < PT_QH; ++j) ihi[j] = imid[j] = ilo[j] = 0;
#pragma unroll
    for (int l = 0; l < kLimbs; ++l) {
      uint32_t v[32];
      tmem_ld_32x32(tmem_base + ((uint32_t)(qw * 32) << 16) + (uint32_t)(l * kLimbQ + half * PT_QH), v);
#pragma unroll
      for (int j = 0; j < PT_QH; ++j) {
        const int iv = __float_as_int(fmaf(__uint_as_float(v[j]), 64.0f, 12582912.0f)) - 0x4B400000;   // planes carry 2^-6; exact float -> int
        if (l < 4) ihi[j] = ihi[j] * 30 + iv;
        else if (l < 8) imid[j] = imid[j] * 30 + iv;
        else ilo[j] = iv;
      }
    }
    double xr[kMaxCov];
    for (int c = 0; c < a.C; ++c) xr[c] = a.xy[(int64_t)t * a.cpp + c];
    const double w3 = 1.0 / 27000.0, w7 = 1.0 / 21870000000.0, w8 = 1.0 / 656100000000.0;   // 30^-3, 30^-7, 30^-8
#pragma unroll
    for (int j = 0; j < PT_QH; ++j) {
      const int qq = half * PT_QH + j;
      if (qq < nq) {
        const int q = q0 + qq;
        const int r = q / a.P, p = q % a.P;
        double val = fma((double)ihi[j], w3, fma((double)imid[j], w7, (double)ilo[j] * w8)) * s_scale[qq];
        for (int c = 0; c < a.C; ++c) val -= xr[c] * s_cvec[qq * a.C + c];
        val *= (double)a.mask[(int64_t)p * a.npad + t];
        a.W[p][(int64_t)(a.col0 + r) * a.npad + t] = val;
      }
    }
    if (a.dbg && threadIdx.x == 64) {
      long long* d = a.dbg + ((int64_t)blockIdx.y * gridDim.x + blockIdx.x) * 4;
      d[0] = t_start; d[1] = t_mma; d[2] = clock64(); d[3] = 0;
    }
  }
  fence_before();
  __syncthreads();
  if (warp == 1) {
    fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// =============================================================================================
// INT8 variant (default).  The operand planes' bytes are 8 x dosage as int8 (bed_expand_fp8_kernel), the coefficients are
// split into FIVE balanced radix-254 digits
//   gamma[i,q] = (s_q / 127) * sum_l d_l[i,q] 254^-l,   d_l in {-127..127}  (int8),
// so tcgen05.mma kind::i8 accumulates exact int32 sums (|sum| <= K2 * 16 * 127 < 2^24 for K2 <= 4096) and the FP64 epilogue
// reassembles the prediction to 254^-5 / 2 = 4.7e-13 s_q per coefficient.  Against the e4m3 version above: 250 instead of 504
// digit rows per pass (5 limbs x 8 bits instead of 9 x 4.9 bits) - half the B-operand traffic from L2, which bounds this
// kernel, half the tensor work, 256 TMEM columns and 96 KiB of shared memory, so TWO CTAs share an SM and one's epilogue
// overlaps the other's main loop.
namespace {
constexpr int PI_BN = 256;
constexpr int PI_A_BYTES = PT_BK * PT_BM;          // 16 KiB
constexpr int PI_B_BYTES = PI_BN * PT_BK;          // 32 KiB
constexpr int PI_QH = kLimbQI8 / 2;                // outputs per epilogue thread (25)
static_assert(kLimbsI8 * kLimbQI8 <= PI_BN && kLimbQI8 % 2 == 0 && PI_QH <= 32 && kLimbQI8 >= 32, "INT8 prediction layout");

// kind::i8, A = B = signed 8-bit (format 1), D = S32 (2), A MN-major (bit 15), B K-major, M = 128, N = 256
constexpr uint32_t kPiIdesc = (2u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | ((uint32_t)(PI_BN >> 3) << 17) |
                              ((uint32_t)(128 >> 4) << 24);

__device__ __forceinline__ void mma_i8(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(da), "l"(db), "r"(kPiIdesc), "r"(accumulate) : "memory");
}
}  // namespace

// digit rows  dig[f][g][l*50 + qq][k]  (int8), scales s[f][q].  grid: (Qp, K folds), block 256.
__global__ void __launch_bounds__(256)
l0_gamma_limbs_i8_kernel(const double* __restrict__ gam, const double* __restrict__ gmu, int Qp, int Q, int bs,
                         int rows_p, double* __restrict__ scale, uint8_t* __restrict__ dig, int ngroups) {
  __shared__ double red[256];
  const int q = blockIdx.x, f = blockIdx.y;
  if (q >= Q) return;
  const double* gcol = gam + (int64_t)f * rows_p * Qp + q;
  const double* mcol = gmu + (int64_t)f * rows_p * Qp + q;
  double mx = 0.0;
  for (int i = threadIdx.x; i < bs; i += 256)
    mx = fmax(mx, fmax(fabs(gcol[(int64_t)i * Qp]), fabs(mcol[(int64_t)i * Qp])));
  red[threadIdx.x] = mx;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] = fmax(red[threadIdx.x], red[threadIdx.x + o]);
    __syncthreads();
  }
  const double s = red[0] > 0.0 ? red[0] : 1.0;
  if (threadIdx.x == 0) scale[(int64_t)f * Qp + q] = s;
  const int g = q / kLimbQI8, qq = q % kLimbQI8;
  const int K2 = 2 * rows_p;
  uint8_t* base = dig + ((int64_t)f * ngroups + g) * (int64_t)PI_BN * K2;
  for (int k = threadIdx.x; k < K2; k += 256) {
    const int plane = k >= rows_p, i = plane ? k - rows_p : k;
    double v = 0.0;
    if (i < bs) v = (plane ? mcol[(int64_t)i * Qp] : gcol[(int64_t)i * Qp]) / s * 127.0;
#pragma unroll
    for (int l = 0; l < kLimbsI8; ++l) {
      const double d = rint(v);                    // |v| <= 127: the remainder (<= 1/2) x 254 stays in range
      base[(int64_t)(l * kLimbQI8 + qq) * K2 + k] = (uint8_t)(int8_t)(int)d;
      v = (v - d) * 254.0;
    }
  }
}

// MT = sample tiles (of 128) per CTA.  MT = 1 (default): 2 stages of 48 KiB, 256 TMEM columns, two CTAs per SM.  MT = 2
// (RG_B200_PREDICT_MT=2): both tiles share every digit-row stage - the B operand, 2/3 of this kernel's L2 -> shared-memory
// traffic, is read once per 256 samples - 3 stages of 64 KiB, all 512 TMEM columns, one CTA per SM.  Measured SLOWER
// (profiles/ab_r2j_ablation.txt: 12.4 vs 10.0 ms per 50 blocks alone, 44.5 vs 42.6 ms overlapped): two co-resident CTAs hide
// the epilogue and the TMA latency better than the saved traffic pays.  (The fold padding is a multiple of 256 samples, so
// the two tiles of a CTA always belong to the same fold.)
// grid: (Npad / (128 MT) sample tiles, q groups); 320 threads.
template <int MT>
__global__ void __launch_bounds__(PT_THREADS, MT == 1 ? 2 : 1)
l0_predict_i8_kernel(const __grid_constant__ CUtensorMap tmZ, const __grid_constant__ CUtensorMap tmD, PredictTcArgs a) {
  constexpr int NST = MT == 1 ? 2 : 3;
  constexpr int STAGE_BYTES = MT * PI_A_BYTES + PI_B_BYTES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gen_base = smem_raw + (base - raw);
  const uint32_t sA = base;                                   // [NST][MT][16 KiB]
  const uint32_t sB = base + NST * MT * PI_A_BYTES;           // [NST][32 KiB]
  uint64_t* bars = reinterpret_cast<uint64_t*>(gen_base + NST * STAGE_BYTES);
  const uint32_t full_bar = smem_u32(bars);
  const uint32_t empty_bar = smem_u32(bars + NST);
  const uint32_t tmem_full_bar = smem_u32(bars + 2 * NST);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * NST + 1);
  double* s_scale = reinterpret_cast<double*>(bars + 2 * NST + 2);         // [kLimbQI8]
  double* s_cvec = s_scale + kLimbQI8;                                      // [kLimbQI8][C]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = blockIdx.x, g = blockIdx.y;
  const int f = a.tile_fold[tile * MT];
  const int nkb = (2 * a.rows_p) / PT_BK;
  const int q0 = g * kLimbQI8;
  const int nq = min(kLimbQI8, a.Q - q0);

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < NST; ++s) { mbar_init(full_bar + 8 * s, 1); mbar_init(empty_bar + 8 * s, 1); }
    mbar_init(tmem_full_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmZ) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmD) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)(MT * PI_BN)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int e = threadIdx.x; e < kLimbQI8; e += PT_THREADS)
    s_scale[e] = (e < nq) ? a.scale[(int64_t)f * a.Qp + q0 + e] / 127.0 : 0.0;
  for (int e = threadIdx.x; e < kLimbQI8 * a.C; e += PT_THREADS) {
    const int qq = e / a.C, c = e % a.C;
    s_cvec[e] = (qq < nq) ? a.cvec[((int64_t)f * a.Qp + q0 + qq) * a.C + c] : 0.0;
  }
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer: A = MT x (128 plane rows x 128 samples); B = 256 digit rows x 128 k bytes =====
      const int drow0 = (f * a.ngroups + g) * PI_BN;
      // the genotype planes stream from DRAM (206 MB per block at N = 100k): their boxes are prefetched into L2 `pf` k-blocks
      // ahead of the shared-memory ring (the digit rows are L2-resident anyway)
      const int pf = a.l2_prefetch;
      auto prefetch = [&](int kb) {
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) tma_prefetch_2d(&tmZ, (tile * MT + mt) * PT_BM, kb * PT_BK);
      };
      for (int kb = NST; kb < NST + pf && kb < nkb; ++kb) prefetch(kb);
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % NST;
        const uint32_t ph = (kb / NST) & 1;
        if (pf > 0 && kb + NST + pf < nkb) prefetch(kb + NST + pf);
        mbar_wait(empty_bar + 8 * s, ph ^ 1);
        mbar_expect_tx(full_bar + 8 * s, STAGE_BYTES);
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
          tma_load_2d(sA + (s * MT + mt) * PI_A_BYTES, &tmZ, full_bar + 8 * s, (tile * MT + mt) * PT_BM, kb * PT_BK);
        tma_load_2d(sB + s * PI_B_BYTES, &tmD, full_bar + 8 * s, kb * PT_BK, drow0);
        tma_load_2d(sB + s * PI_B_BYTES + 16384, &tmD, full_bar + 8 * s, kb * PT_BK, drow0 + 128);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===== MMA issuer =====
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % NST;
        const uint32_t ph = (kb / NST) & 1;
        mbar_wait(full_bar + 8 * s, ph);
        fence_after();
        const uint64_t db = make_desc(sB + s * PI_B_BYTES);
#pragma unroll
        for (int k = 0; k < PT_BK / 32; ++k) {
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) {
            const uint64_t da = make_desc(sA + (s * MT + mt) * PI_A_BYTES);
            mma_i8(tmem_base + (uint32_t)(mt * PI_BN), da + (uint64_t)(256 * k), db + (uint64_t)(2 * k), (kb | k) ? 1u : 0u);
          }
        }
        tcgen05_commit(empty_bar + 8 * s);
      }
      tcgen05_commit(tmem_full_bar);
    }
  } else {
    // ===== epilogue: thread = (sample, half of the outputs).  Limb sums are exact int32 multiples of 8; FP64 Horner from the
    // lowest limb up.  The second half reads the 32 columns that END at its last output, so no load leaves the allocation.
    const int qw = warp & 3;                      // TMEM lane quarter this warp may access
    const int half = (warp - 2) >> 2;
    mbar_wait(tmem_full_bar, 0);
    fence_after();
    const double inv254 = 1.0 / 254.0;
#pragma unroll 1
    for (int mt = 0; mt < MT; ++mt) {
      const int t = (tile * MT + mt) * PT_BM + qw * 32 + lane;
      const uint32_t tbase = tmem_base + ((uint32_t)(qw * 32) << 16) + (uint32_t)(mt * PI_BN);
      double acc[PI_QH];
#pragma unroll
      for (int j = 0; j < PI_QH; ++j) acc[j] = 0.0;
      if (half == 0) {
#pragma unroll
        for (int l = kLimbsI8 - 1; l >= 0; --l) {
          uint32_t v[32];
          tmem_ld_32x32(tbase + (uint32_t)(l * kLimbQI8), v);
#pragma unroll
          for (int j = 0; j < PI_QH; ++j) acc[j] = fma(acc[j], inv254, (double)((int)v[j] >> 3));
        }
      } else {
#pragma unroll
        for (int l = kLimbsI8 - 1; l >= 0; --l) {
          uint32_t v[32];
          tmem_ld_32x32(tbase + (uint32_t)(l * kLimbQI8 + kLimbQI8 - 32), v);
#pragma unroll
          for (int j = 0; j < PI_QH; ++j) acc[j] = fma(acc[j], inv254, (double)((int)v[32 - PI_QH + j] >> 3));
        }
      }
      double xr[kMaxCov];
      for (int c = 0; c < a.C; ++c) xr[c] = a.xy[(int64_t)t * a.cpp + c];
#pragma unroll
      for (int j = 0; j < PI_QH; ++j) {
        const int qq = half * PI_QH + j;
        if (qq < nq) {
          const int q = q0 + qq;
          const int r = q / a.P, p = q % a.P;
          double val = acc[j] * s_scale[qq];
          for (int c = 0; c < a.C; ++c) val -= xr[c] * s_cvec[qq * a.C + c];
          val *= (double)a.mask[(int64_t)p * a.npad + t];
          a.W[p][(int64_t)(a.col0 + r) * a.npad + t] = val;
        }
      }
    }
  }
  fence_before();
  __syncthreads();
  if (warp == 1) {
    fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)(MT * PI_BN)) : "memory");
  }
}

// Column sums of the raw predictions for the standardisation: part[chunk][q] = (sum, sum of squares) over
// a chunk of 8192 samples, fixed-order tree reduction.  grid: (Q, nchunks), block 256.
__global__ void __launch_bounds__(256)
l0_colsum_kernel(double* const* __restrict__ W, int64_t npad, int col0, int P, int Qp,
                 double* __restrict__ part) {
  __shared__ double r1[256], r2[256];
  const int q = blockIdx.x, r = q / P, p = q % P;
  const int64_t t0 = (int64_t)blockIdx.y * 8192;
  const double* w = W[p] + (int64_t)(col0 + r) * npad;
  double s1 = 0.0, s2 = 0.0;
  for (int64_t t = t0 + threadIdx.x; t < min(t0 + 8192, npad); t += 256) {
    const double v = w[t];
    s1 += v;
    s2 = fma(v, v, s2);
  }
  r1[threadIdx.x] = s1; r2[threadIdx.x] = s2;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) { r1[threadIdx.x] += r1[threadIdx.x + o]; r2[threadIdx.x] += r2[threadIdx.x + o]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    part[((int64_t)blockIdx.y * Qp + q) * 2 + 0] = r1[0];
    part[((int64_t)blockIdx.y * Qp + q) * 2 + 1] = r2[0];
  }
}

int launch_l0_colsum(double* const* W, int64_t npad, int col0, int P, int Q, int Qp, double* part,
                     cudaStream_t s) {
  const int nchunks = (int)ceil_div(npad, 8192);
  dim3 grid(Q, nchunks);
  l0_colsum_kernel<<<grid, 256, 0, s>>>(W, npad, col0, P, Qp, part);
  return nchunks;
}

// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn pt_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    RG_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres));
    RG_CHECK(p != nullptr && qres == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled not available");
    fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// 2D byte tensor [rows][inner] with a 128 x 128 box and 128B swizzle
void make_byte_tensor_map(CUtensorMap* tm, const uint8_t* basep, int64_t inner, int64_t rows) {
  const cuuint64_t gdim[2] = {(cuuint64_t)inner, (cuuint64_t)rows};
  const cuuint64_t gstride[1] = {(cuuint64_t)inner};
  const cuuint32_t box[2] = {128, 128};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = pt_encode_fn()(tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<uint8_t*>(basep), gdim, gstride, box,
                              estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  RG_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")");
}

size_t predict_tc_dig_bytes(int K, int ngroups, int rows_p) { return (size_t)K * ngroups * PT_BN * 2 * rows_p; }

void launch_l0_gamma_limbs(const double* gam, const double* gmu, int Qp, int Q, int bs, int rows_p, int K,
                           double* scale, uint8_t* dig, int ngroups, cudaStream_t s) {
  dim3 grid(Qp, K);
  l0_gamma_limbs_kernel<<<grid, 256, 0, s>>>(gam, gmu, Qp, Q, bs, rows_p, scale, dig, ngroups);
}

void launch_l0_predict_tcgen05(const CUtensorMap& tmZ, const CUtensorMap& tmD, const PredictTcArgs& a, int ntiles,
                               cudaStream_t s) {
  const size_t smem = (size_t)PT_STAGES * PT_STAGE_BYTES + 1024 + 128 +
                      ((size_t)kLimbQ * (1 + a.C)) * sizeof(double);
  ensure_dyn_smem(reinterpret_cast<const void*>(l0_predict_tcgen05_kernel), smem);
  dim3 grid(ntiles, a.ngroups);
  l0_predict_tcgen05_kernel<<<grid, PT_THREADS, smem, s>>>(tmZ, tmD, a);
}

size_t predict_i8_dig_bytes(int K, int ngroups, int rows_p) { return (size_t)K * ngroups * PI_BN * 2 * rows_p; }

void launch_l0_gamma_limbs_i8(const double* gam, const double* gmu, int Qp, int Q, int bs, int rows_p, int K,
                              double* scale, uint8_t* dig, int ngroups, cudaStream_t s) {
  dim3 grid(Qp, K);
  l0_gamma_limbs_i8_kernel<<<grid, 256, 0, s>>>(gam, gmu, Qp, Q, bs, rows_p, scale, dig, ngroups);
}

void launch_l0_predict_i8(const CUtensorMap& tmZ, const CUtensorMap& tmD, const PredictTcArgs& a, int ntiles,
                          cudaStream_t s) {
  RG_CHECK(2 * a.rows_p <= 4096, "INT8 prediction: 2 * rows_p <= 4096 (int32 Horner bound)");
  static const int mt = [] { const char* e = getenv("RG_B200_PREDICT_MT"); return (e && atoi(e) == 2) ? 2 : 1; }();
  const size_t extra = 1024 + 128 + ((size_t)kLimbQI8 * (1 + a.C)) * sizeof(double);
  if (mt == 2 && ntiles % 2 == 0) {
    const size_t smem = (size_t)3 * (2 * PI_A_BYTES + PI_B_BYTES) + extra;
    ensure_dyn_smem(reinterpret_cast<const void*>(l0_predict_i8_kernel<2>), smem);
    l0_predict_i8_kernel<2><<<dim3(ntiles / 2, a.ngroups), PT_THREADS, smem, s>>>(tmZ, tmD, a);
  } else {
    const size_t smem = (size_t)2 * (PI_A_BYTES + PI_B_BYTES) + extra;
    ensure_dyn_smem(reinterpret_cast<const void*>(l0_predict_i8_kernel<1>), smem);
    l0_predict_i8_kernel<1><<<dim3(ntiles, a.ngroups), PT_THREADS, smem, s>>>(tmZ, tmD, a);
  }
}

}  // namespace rg
