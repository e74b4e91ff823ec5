// Per-fold integer Gram  Z_f Z_f^T  on the 5th-generation tensor cores (tcgen05 + TMEM + TMA).
//
// Z = [G0; Miss] is the (2*rows_p) x Npad e4m3 operand written by bed_expand_fp8_kernel:
// G0 in {0,1,2} with missing calls as 0, Miss in {0,1}.  Every product is an integer <= 4 and
// every fold sum is < 2^24, so the FP32 accumulators in TMEM hold the EXACT integer Grams
//   G0 G0^T, Miss G0^T, Miss Miss^T   restricted to the fold's sample range,
// which is all of Data::calc_cv_matrices' bs x bs x N work (reference src/Data.cpp:748:
// `G_folds[i] = Gmat * Gmat.transpose()`, 2*bs^2*N flops) - the rank-C covariate/scale/mean
// corrections are applied afterwards in FP64 (l0_stats.cu).
//
// Kernel shape: one CTA per (128 x 256) output tile of the lower triangle per fold.
//   warp 0      : TMA producer  (cp.async.bulk.tensor 2D, 128B swizzle, 4-stage mbarrier ring)
//   warp 1      : TMEM allocator + single-thread tcgen05.mma issuer (kind::f8f6f4, M128 N256 K32)
//   warps 2..5  : epilogue, tcgen05.ld 32x32b -> registers -> coalesced 128-bit global stores
// K loop = the fold's samples in 128-byte (= 128-sample) swizzle atoms, 4 MMAs per atom.
#include <stdlib.h>

#include "kernels.cuh"

namespace rg {

namespace {

constexpr int BM = 128;
constexpr int BK = 128;             // bytes == samples per K step (one 128B swizzle atom)
constexpr int STAGES = 4;
constexpr int A_BYTES = BM * BK;    // 16 KiB
// BN (template parameter): 256 for the Gram and wide statistics tiles (B stage 32 KiB), 128 for a single 128-row digit group
constexpr int NTHREADS = 192;
constexpr uint32_t SPIN_LIMIT = 1u << 28;   // bounded waits: a protocol bug traps instead of hanging

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
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
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) break;
    if (++spins > SPIN_LIMIT) __trap();
  }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(tm), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tcgen05_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major, 128B-swizzled shared-memory matrix descriptor (8-row groups 1024 B apart).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t addr) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);   // start address  [0,14)
  d |= (uint64_t)1 << 16;                  // LBO (unused for swizzled K-major) [16,30)
  d |= (uint64_t)(1024 >> 4) << 32;        // SBO = 1024 B    [32,46)
  d |= (uint64_t)1 << 46;                  // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                  // SWIZZLE_128B
  return d;
}

// kind::f8f6f4, A = B = E4M3 (format 0), D = F32 (1), both K-major, M = 128, N = bn.
__host__ __device__ constexpr uint32_t gram_idesc(int bn) { return (1u << 4) | ((uint32_t)(bn >> 3) << 17) | ((uint32_t)(BM >> 4) << 24); }

__device__ __forceinline__ void mma_f8(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t accumulate, uint32_t idesc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
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

}  // namespace

// grid: (ntiles, K folds)
template <int BN>
__global__ void __launch_bounds__(NTHREADS, 1)
gram_fp8_tcgen05_kernel(const __grid_constant__ CUtensorMap tmZ, const __grid_constant__ CUtensorMap tmB,
                        const int2* __restrict__ tiles,
                        const int2* __restrict__ fold_k, float* __restrict__ out, int ldo,
                        int64_t fold_stride, float out_scale) {
  constexpr int B_BYTES = BN * BK;
  constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr int TMEM_COLS = BN;
  extern __shared__ uint8_t smem_raw[];
  // 128B swizzle needs 1024-byte aligned stage buffers
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gen_base = smem_raw + (base - raw);
  const uint32_t sA = base;
  const uint32_t sB = base + STAGES * A_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(gen_base + STAGES * STAGE_BYTES);
  const uint32_t full_bar = smem_u32(bars);                  // [STAGES]
  const uint32_t empty_bar = smem_u32(bars + STAGES);        // [STAGES]
  const uint32_t tmem_full_bar = smem_u32(bars + 2 * STAGES);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int2 tile = tiles[blockIdx.x];          // (m tile of 128 rows, n tile of 256 rows)
  const int2 fk = fold_k[blockIdx.y];           // (first K block, number of K blocks)
  const int nkb = fk.y;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar + 8 * s, 1);
      mbar_init(empty_bar + 8 * s, 1);
    }
    mbar_init(tmem_full_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmZ) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer =====
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(empty_bar + 8 * s, ph ^ 1);
        mbar_expect_tx(full_bar + 8 * s, STAGE_BYTES);
        const int kc = (fk.x + kb) * BK;
        tma_load_2d(sA + s * A_BYTES, &tmZ, full_bar + 8 * s, kc, tile.x * BM);
        tma_load_2d(sB + s * B_BYTES, &tmB, full_bar + 8 * s, kc, tile.y * BN);
        if (BN == 256) tma_load_2d(sB + s * B_BYTES + A_BYTES, &tmB, full_bar + 8 * s, kc, tile.y * BN + 128);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===== MMA issuer (one thread) =====
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(full_bar + 8 * s, ph);
        tcgen05_fence_after();
        const uint64_t da = make_smem_desc(sA + s * A_BYTES);
        const uint64_t db = make_smem_desc(sB + s * B_BYTES);
#pragma unroll
        for (int k = 0; k < BK / 32; ++k) {
          // advance 32 bytes (= K of one f8 MMA) inside the swizzle atom: +2 in 16-byte units
          mma_f8(tmem_base, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), (kb | k) ? 1u : 0u, gram_idesc(BN));
        }
        tcgen05_commit(empty_bar + 8 * s);     // frees the smem stage when these MMAs retire
      }
      tcgen05_commit(tmem_full_bar);           // accumulator complete
    }
  } else {
    // ===== epilogue: TMEM -> registers -> global =====
    const int q = warp & 3;                    // TMEM lane quarter this warp may access
    mbar_wait(tmem_full_bar, 0);
    tcgen05_fence_after();
    const int row = tile.x * BM + q * 32 + lane;
    float* orow = out + (int64_t)blockIdx.y * fold_stride + (int64_t)row * ldo + tile.y * BN;
#pragma unroll 1
    for (int c = 0; c < BN / 32; ++c) {
      uint32_t v[32];
      tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32), v);
      float4* o4 = reinterpret_cast<float4*>(orow + c * 32);
#pragma unroll
      for (int j = 0; j < 8; ++j)
        o4[j] = make_float4(__uint_as_float(v[4 * j]) * out_scale, __uint_as_float(v[4 * j + 1]) * out_scale,
                            __uint_as_float(v[4 * j + 2]) * out_scale, __uint_as_float(v[4 * j + 3]) * out_scale);
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS)
                 : "memory");
  }
}

// ---------------------------------------------------------------------------------------
// Test-only reference: same quantity on CUDA cores straight from the 2-bit codes (used by
// tests through rg_debug_fetch to localise a tensor-core protocol bug; never on the product path).
__global__ void gram_reference_kernel(const uint8_t* __restrict__ z, int64_t npad, int rows2, int k0, int k1,
                                      float* __restrict__ out, int ldo) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int i = blockIdx.y;
  if (j >= rows2 || j > i) return;
  const uint8_t* zi = z + (int64_t)i * npad;
  const uint8_t* zj = z + (int64_t)j * npad;
  int acc = 0;
  for (int t = k0; t < k1; ++t) {
    const int a = zi[t] >> 3;        // plane bytes 0x00 / 0x08 / 0x10 (bed_expand_fp8_kernel)
    const int b = zj[t] >> 3;
    acc += a * b;
  }
  out[(int64_t)i * ldo + j] = (float)acc;
}

// ---------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
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

void make_gram_tensor_map(CUtensorMap* tm, const uint8_t* z, int64_t npad, int rows2) {
  const cuuint64_t gdim[2] = {(cuuint64_t)npad, (cuuint64_t)rows2};
  const cuuint64_t gstride[1] = {(cuuint64_t)npad};
  const cuuint32_t box[2] = {BK, 128};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = get_encode_fn()(tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<uint8_t*>(z), gdim, gstride, box,
                               estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                               CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  RG_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")");
}

size_t gram_smem_bytes(int bn) {
  static const bool exclusive = getenv("RG_DBG_GRAM_EXCLUSIVE") != nullptr;
  return exclusive ? (size_t)232448 : (size_t)STAGES * (A_BYTES + bn * BK) + 1024 + 128;
}

void gram_tile_list(int rows2, std::vector<int2>& tiles) {
  tiles.clear();
  for (int nj = 0; nj < rows2 / 256; ++nj)
    for (int mi = 2 * nj; mi < rows2 / BM; ++mi) tiles.push_back(make_int2(mi, nj));
}

void launch_gram_tcgen05(const CUtensorMap& tm, const CUtensorMap& tmB, const int2* tiles, int ntiles, const int2* fold_k, int K,
                         float* out, int ldo, int64_t fold_stride, float out_scale, cudaStream_t s, int bn) {
  RG_CHECK(bn == 256 || bn == 128, "gram tiles are 128 x 256 or 128 x 128");
  dim3 grid(ntiles, K);
  if (bn == 256) {
    ensure_dyn_smem(reinterpret_cast<const void*>(gram_fp8_tcgen05_kernel<256>), gram_smem_bytes(256));
    gram_fp8_tcgen05_kernel<256><<<grid, NTHREADS, gram_smem_bytes(256), s>>>(tm, tmB, tiles, fold_k, out, ldo, fold_stride, out_scale);
  } else {
    ensure_dyn_smem(reinterpret_cast<const void*>(gram_fp8_tcgen05_kernel<128>), gram_smem_bytes(128));
    gram_fp8_tcgen05_kernel<128><<<grid, NTHREADS, gram_smem_bytes(128), s>>>(tm, tmB, tiles, fold_k, out, ldo, fold_stride, out_scale);
  }
}

void launch_gram_reference(const uint8_t* z, int64_t npad, int rows2, int k0, int k1, float* out, int ldo,
                           cudaStream_t s) {
  dim3 grid((unsigned)ceil_div(rows2, 128), rows2);
  gram_reference_kernel<<<grid, 128, 0, s>>>(z, npad, rows2, k0, k1, out, ldo);
}

}  // namespace rg
