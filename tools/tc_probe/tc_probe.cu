// Development probe (not part of the library): one tcgen05.mma tile, D[64 x N] = A[64 x K] * B[N x K]^T in 3xTF32
// (hi/lo split, fp32-level accuracy), operands in the canonical no-swizzle K-major layout [K/4][rows][4], accumulator in
// TMEM, read back with tcgen05.ld.  Validates descriptor encodings and the TMEM row mapping before the scorer uses them.
#include <cuda_runtime.h>
#include <stdint.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// shared-memory matrix descriptor: no swizzle, K-major, version 1 (cute/arch/mma_sm100_desc.hpp SmemDescriptor)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // version
  return d;
}
// instruction descriptor: F32 accumulate, TF32 x TF32, both K-major, M x N
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

template <int N, int K>
__global__ void __launch_bounds__(128) tc_probe_kernel(const float* A, const float* B, float* D, int mode) {
  extern __shared__ __align__(128) unsigned char smem[];
  float* a_hi = (float*)smem;                    // [K/4][64][4]
  float* a_lo = a_hi + 64 * K;
  float* b_hi = a_lo + 64 * K;                   // [K/4][N][4]
  float* b_lo = b_hi + N * K;
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int idx = tid; idx < 64 * K; idx += 128) {
    const int m = idx / K, k = idx % K;
    const float x = A[idx];
    const float h = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
    const int o = ((k >> 2) * 64 + m) * 4 + (k & 3);
    a_hi[o] = h;
    a_lo[o] = x - h;
  }
  for (int idx = tid; idx < N * K; idx += 128) {
    const int n = idx / K, k = idx % K;
    const float x = B[idx];
    const float h = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
    const int o = ((k >> 2) * N + n) * 4 + (k & 3);
    b_hi[o] = h;
    b_lo[o] = x - h;
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy smem writes -> visible to the tensor core
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "n"(64));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t taddr = tmem_base;
  if (tid == 0) {
    const uint32_t idesc = make_idesc(64, N);
    const uint32_t lbo_a = 64 * 16, lbo_b = N * 16, sbo = 128;
    int first = 1;
    const int passes = mode == 0 ? 1 : 3;
    for (int p = 0; p < passes; ++p) {
      const float* pa = p == 2 ? a_lo : a_hi;
      const float* pb = p == 1 ? b_lo : b_hi;
      for (int k8 = 0; k8 < K / 8; ++k8) {
        const uint64_t da = make_desc(smem_u32(pa) + k8 * 2 * lbo_a, lbo_a, sbo);
        const uint64_t db = make_desc(smem_u32(pb) + k8 * 2 * lbo_b, lbo_b, sbo);
        const uint32_t acc = first ? 0u : 1u;
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(taddr),
            "l"(da), "l"(db), "r"(idesc), "r"(acc)
            : "memory");
        first = 0;
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  // wait for the MMAs
  {
    uint32_t done = 0;
    const uint32_t addr = smem_u32(&bar);
    do {
      asm volatile(
          "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
          : "=r"(done)
          : "r"(addr), "r"(0u)
          : "memory");
    } while (!done);
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  // M = 64, cta_group::1: row m lives in TMEM lane (m % 16) + 32 * (m / 16); warp w reads the lanes of subpartition w
  for (int c0 = 0; c0 < N; c0 += 16) {
    uint32_t r[16];
    const uint32_t a = taddr + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(a));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    // every lane stores what it got: the host works the mapping out (lane l of warp w -> output row l + 32 w of a
    // [128][N] dump)
    for (int j = 0; j < 16; ++j) D[(size_t)(warp * 32 + lane) * N + c0 + j] = __uint_as_float(r[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(64));
}

extern "C" int tc_probe(const float* dA, const float* dB, float* dD, int N, int K, int mode) {
  if (N != 64 || K != 64) return -22;
  const size_t smem = (size_t)(2 * 64 * K + 2 * N * K) * sizeof(float);
  cudaFuncSetAttribute(tc_probe_kernel<64, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  tc_probe_kernel<64, 64><<<1, 128, smem>>>(dA, dB, dD, mode);
  cudaError_t e = cudaDeviceSynchronize();
  return e == cudaSuccess ? 0 : -1000 - (int)e;
}
