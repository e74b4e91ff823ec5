// Probe: tcgen05.mma kind::f8f6f4 with an MN-major A operand; find the descriptor convention empirically.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/mma_probe tools/mma_probe.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t mkdesc(uint32_t addr, uint32_t lbo16, uint32_t sbo16) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)(lbo16 & 0x3FFF) << 16;
  d |= (uint64_t)(sbo16 & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// A_mn: smem tile rows = k (KT rows) x 128 bytes of m, 128B-swizzled (as TMA would write a Z tile)
// B   : smem tile rows = n (256) x 128 bytes of k, 128B-swizzled K-major (only first 32..KT bytes used)
__global__ void probe(const uint8_t* Amn /*[KT][128]*/, const uint8_t* Bk /*[256][128]*/, float* D /*[128][256]*/,
                      int KT, uint32_t lbo16, uint32_t sbo16, uint32_t a_step16, int a_major) {
  extern __shared__ uint8_t raw[];
  const uint32_t r = smem_u32(raw);
  const uint32_t base = (r + 1023u) & ~1023u;
  uint8_t* g = raw + (base - r);
  uint8_t* sA = g;                 // KT*128 bytes (<= 16 KB)
  uint8_t* sB = g + 16384;         // 32 KB
  uint64_t* bar = reinterpret_cast<uint64_t*>(g + 16384 + 32768);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  // swizzled fill: byte (row, col) -> row*128 + ((col/16) ^ (row%8))*16 + col%16
  for (int e = threadIdx.x; e < KT * 128; e += blockDim.x) {
    const int row = e / 128, col = e % 128;
    sA[row * 128 + (((col >> 4) ^ (row & 7)) << 4) + (col & 15)] = Amn[e];
  }
  for (int e = threadIdx.x; e < 256 * 128; e += blockDim.x) {
    const int row = e / 128, col = e % 128;
    sB[row * 128 + (((col >> 4) ^ (row & 7)) << 4) + (col & 15)] = Bk[e];
  }
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy smem writes -> async proxy (MMA)
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(256u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = *slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = (1u << 4) | ((uint32_t)a_major << 15) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    for (int k = 0; k < KT / 32; ++k) {
      const uint64_t da = mkdesc(smem_u32(sA), lbo16, sbo16) + (uint64_t)(a_step16 * k);
      const uint64_t db = mkdesc(smem_u32(sB), 1, 64) + (uint64_t)(2 * k);
      const uint32_t acc = k ? 1u : 0u;
      asm volatile(
          "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
          "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}" ::"r"(tm), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
  }
  // wait
  {
    uint32_t done = 0; int spins = 0;
    while (!done) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(done) : "r"(smem_u32(bar)), "r"(0u) : "memory");
      if (++spins > (1 << 26)) { __trap(); }
    }
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (threadIdx.x < 128) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int c = 0; c < 8; ++c) {
      uint32_t v[32];
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
          "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
          "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
          : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
            "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
            "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
            "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
          : "r"(tm + ((uint32_t)(warp * 32) << 16) + (uint32_t)(c * 32)));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int j = 0; j < 32; ++j) D[(warp * 32 + lane) * 256 + c * 32 + j] = __uint_as_float(v[j]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(256u) : "memory");
}

static const uint8_t E[16] = {0x00, 0x38, 0x40, 0x44, 0x48, 0x4A, 0x4C, 0x4E, 0x50, 0x51, 0x52, 0x53, 0x54, 0x55, 0x56, 0x57};

int main() {
  const int KT = 128;
  std::vector<int> Ai(128 * KT), Bi(256 * KT);
  srand(1);
  for (auto& v : Ai) v = rand() % 3;          // genotype-like 0..2
  for (auto& v : Bi) v = rand() % 31 - 15;    // digits -15..15
  std::vector<uint8_t> Amn(KT * 128), Bk(256 * 128, 0);
  for (int k = 0; k < KT; ++k) for (int m = 0; m < 128; ++m) Amn[k * 128 + m] = E[Ai[m * KT + k]];
  for (int n = 0; n < 256; ++n) for (int k = 0; k < KT; ++k) { int d = Bi[n * KT + k]; Bk[n * 128 + k] = E[d < 0 ? -d : d] | (d < 0 ? 0x80 : 0); }
  uint8_t *dA, *dB; float* dD;
  cudaMalloc(&dA, Amn.size()); cudaMalloc(&dB, Bk.size()); cudaMalloc(&dD, 128 * 256 * 4);
  cudaMemcpy(dA, Amn.data(), Amn.size(), cudaMemcpyHostToDevice);
  cudaMemcpy(dB, Bk.data(), Bk.size(), cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 60000);
  struct V { uint32_t lbo, sbo, step; const char* name; } vs[] = {
      {1, 64, 256, "lbo=1 sbo=64 step=256 (expected)"}, {64, 1, 256, "lbo=64 sbo=1 step=256"},
      {64, 64, 256, "lbo=64 sbo=64 step=256"}, {1, 64, 2, "lbo=1 sbo=64 step=2"},
      {128, 64, 256, "lbo=128 sbo=64"}, {64, 128, 256, "lbo=64 sbo=128"}, {8, 64, 256, "lbo=8 sbo=64"}, {64, 8, 256, "lbo=64 sbo=8"},
      {0, 64, 256, "lbo=0 sbo=64"}, {256, 64, 256, "lbo=256 sbo=64"}};
  for (int kt : {32, 128}) {
    for (auto& v : vs) {
      cudaMemset(dD, 0, 128 * 256 * 4);
      probe<<<1, 256, 60000>>>(dA, dB, dD, kt, v.lbo, v.sbo, v.step, 1);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("KT=%d %s: CUDA error %s\n", kt, v.name, cudaGetErrorString(e)); return 1; }
      std::vector<float> D(128 * 256);
      cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
      long bad = 0; double maxd = 0;
      for (int m = 0; m < 128; ++m) for (int n = 0; n < 256; ++n) {
        int ref = 0; for (int k = 0; k < kt; ++k) ref += Ai[m * KT + k] * Bi[n * KT + k];
        if (D[m * 256 + n] != (float)ref) { ++bad; double dd = fabs(D[m * 256 + n] - ref); if (dd > maxd) maxd = dd; }
      }
      printf("KT=%3d %-36s mismatches=%ld maxdiff=%g  D[0][0..3]=%g %g %g %g\n", kt, v.name, bad, maxd, D[0], D[1], D[2], D[3]);
    }
  }
  return 0;
}
