// Bring-up probe for the 2-CTA tensor-core path (tcgen05 cta_group::2): one 128 x 256 x K GEMM on
// a CTA pair, M = 128 across the pair (64 rows of A and 128 rows of B per CTA), raw TMEM dump of both
// CTAs.  Pins down, on the hardware, everything the paired SIREN kernel relies on: cluster launch,
// cta_group::2 allocation, remote mbarrier arrive, multicast commit, descriptor semantics and the
// accumulator layout (lane = row + 64 * (column >= 128), TMEM column = column % 128).  Test-only
// entry point (tests/test_gpu_siren.py), not used by the operators.
#include "common.cuh"
#include "isob200.h"
#include <cuda_fp16.h>

namespace isob200 {
namespace probe {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  for (uint32_t spin = 0; !done; ++spin) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (spin > (1u << 22)) __trap();
  }
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(128 >> 4) << 32) | (1ull << 46);
}

constexpr int KMAX = 64;

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128)
umma2_probe_kernel(const float* __restrict__ A, const float* __restrict__ B, int K, float* __restrict__ dump) {
  __shared__ __align__(1024) unsigned char sA[64 * KMAX * 2];    // 64 rows of A, K-major core matrices
  __shared__ __align__(1024) unsigned char sB[128 * KMAX * 2];   // 128 rows of B
  __shared__ __align__(8) unsigned long long bars[2];
  __shared__ uint32_t tmem_ptr;
  const uint32_t rank = cluster_rank();
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t bar_ab = smem_u32(&bars[0]), bar_done = smem_u32(&bars[1]);
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar_ab), "r"(2));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar_done), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  cluster_sync();
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_ptr)), "r"(256)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  // operands: fp16, K-major no-swizzle core matrices (8 rows x 16 B); chunk stride = rows * 16
  for (int i = tid; i < 64 * K; i += 128) {
    const int m = i / K, k = i % K;
    const __half v = __float2half_rn(A[(size_t)(64 * rank + m) * K + k]);
    *reinterpret_cast<__half*>(sA + (k / 8) * 1024 + (m / 8) * 128 + (m % 8) * 16 + (k % 8) * 2) = v;
  }
  for (int i = tid; i < 128 * K; i += 128) {
    const int n = i / K, k = i % K;
    const __half v = __float2half_rn(B[(size_t)(128 * rank + n) * K + k]);
    *reinterpret_cast<__half*>(sB + (k / 8) * 2048 + (n / 8) * 128 + (n % 8) * 16 + (k % 8) * 2) = v;
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_ptr;
  if (tid == 0) {
    // operands of this CTA are in place: arrive on the LEADER's barrier (remote for rank 1)
    asm volatile(
        "{\n"
        ".reg .b32 ra;\n"
        "mapa.shared::cluster.u32 ra, %0, %1;\n"
        "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n"
        "}\n" ::"r"(bar_ab),
        "r"(0)
        : "memory");
  }
  if (rank == 0 && tid == 32) {
    mbar_wait(bar_ab, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t idesc = (1u << 4) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    for (int k16 = 0; k16 < K / 16; ++k16) {
      const uint64_t da = make_desc(smem_u32(sA) + k16 * 2 * 1024, 1024);
      const uint64_t db = make_desc(smem_u32(sB) + k16 * 2 * 2048, 2048);
      asm volatile(
          "{\n"
          ".reg .pred p;\n"
          "setp.ne.b32 p, %4, 0;\n"
          "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
          "}\n" ::"r"(tmem_base),
          "l"(da), "l"(db), "r"(idesc), "r"(k16 ? 1u : 0u)
          : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::
                     "r"(bar_done),
                 "h"((unsigned short)3)
                 : "memory");
  }
  mbar_wait(bar_done, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  // raw dump: dump[rank][lane 0..127][column 0..127]
  for (int c0 = 0; c0 < 128; c0 += 32) {
    uint32_t r[32];
    const uint32_t taddr = tmem_base + ((uint32_t)(32 * warp) << 16) + c0;
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int i = 0; i < 32; ++i)
      dump[((size_t)rank * 128 + tid) * 128 + c0 + i] = __uint_as_float(r[i]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  cluster_sync();
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256) : "memory");
}

// Issue-rate microbenchmark: `reps` back-to-back K = 16 MMAs of one shape on (garbage) shared-memory
// operands, cycles from the first issue to the completion of the commit.  mode 0: cta_group::1, M = 128;
// mode 1: cta_group::2, M = 128 (64 rows per CTA); mode 2: cta_group::2, M = 256 (128 rows per CTA).  N = 256.
// mode + 10: the same shapes with 128-byte-swizzled K-major operands (rows of 128 B, 8-row atoms of 1 KB)
// instead of the no-swizzle core-matrix layout the SIREN kernels use.  mode 3: cta_group::1, M = 64.
// mode + 100 n: N = 256 >> n (n = 0..3) instead of 256.  mode + 1000 (cta_group::1 only): consecutive MMAs
// alternate between two accumulators (is the rate a dependent-accumulate latency?); mode + 2000
// (cta_group::1 only): A operand read from tensor memory instead of shared memory.
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128)
umma_rate_kernel(int mode_in, int reps, long long* __restrict__ cycles) {
  __shared__ __align__(1024) unsigned char sB[256 * 128];   // 256 rows x up to 128 B (contents irrelevant:
  unsigned char* sA = sB;                                    //  A reads the same buffer)
  const bool sw128 = (mode_in / 10) % 10 != 0;
  const int mode = mode_in % 10 == 3 ? 0 : mode_in % 10;
  const bool m64 = mode_in % 10 == 3;
  const uint32_t nn = 256u >> ((mode_in / 100) % 10);
  const int variant = mode_in / 1000;   // 0 plain, 1 two accumulators, 2 A from TMEM
  __shared__ __align__(8) unsigned long long bars[1];
  __shared__ uint32_t tmem_ptr;
  const uint32_t rank = cluster_rank();
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t bar_done = smem_u32(&bars[0]);
  for (int i = tid; i < 256 * 128 / 4; i += 128) reinterpret_cast<uint32_t*>(sB)[i] = 0x3c003c00u;
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar_done), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  cluster_sync();
  if (warp == 0) {
    if (mode == 0) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_ptr)), "r"(512) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_ptr)), "r"(256) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  cluster_sync();
  const uint32_t tmem_base = tmem_ptr;
  const bool issuer = tid == 32 && (mode == 0 || rank == 0);
  long long t0 = 0;
  if (issuer) {
    const int a_rows = mode == 1 ? 64 : 128, b_rows = mode == 0 ? 256 : 128;
    const uint32_t m = mode == 2 ? 256 : (m64 ? 64 : 128);
    const uint32_t idesc = (1u << 4) | ((nn >> 3) << 17) | ((m >> 4) << 24);
    // SW128 K-major: SBO = 1024 (8 rows x 128 B), LBO unused, layout type 2 in bits 61..63
    const uint64_t sw = (uint64_t)((1024 >> 4)) << 32 | (1ull << 16) | (1ull << 46) | (2ull << 61);
    const uint64_t da = sw128 ? (sw | ((smem_u32(sA) >> 4) & 0x3FFF)) : make_desc(smem_u32(sA), a_rows * 16);
    const uint64_t db = sw128 ? (sw | ((smem_u32(sB) >> 4) & 0x3FFF)) : make_desc(smem_u32(sB), b_rows * 16);
    t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      if (mode == 0 && variant == 2)
        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n" ::"r"(tmem_base), "r"(tmem_base + 256), "l"(db), "r"(idesc), "r"(r ? 1u : 0u) : "memory");
      else if (mode == 0)
        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_base + ((variant == 1 && (r & 1)) ? 256u : 0u)), "l"(da), "l"(db), "r"(idesc), "r"(r > 1 ? 1u : 0u) : "memory");
      else
        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_base), "l"(da), "l"(db), "r"(idesc), "r"(r ? 1u : 0u) : "memory");
    }
    if (mode == 0)
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_done) : "memory");
    else
      asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar_done), "h"((unsigned short)3) : "memory");
  }
  mbar_wait(bar_done, 0);
  if (issuer) cycles[rank] = clock64() - t0;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  cluster_sync();
  if (warp == 0) {
    if (mode == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256) : "memory");
  }
}

}  // namespace probe
}  // namespace isob200

extern "C" int isob200_umma_rate(int mode, int reps, long long* cycles_dev, void* stream) {
  using namespace isob200;
  ISO_CHECK_ARG(mode >= 0 && mode % 10 <= 3 && (mode / 10) % 10 <= 1 && (mode / 100) % 10 <= 3 && mode < 3000 &&
                    (mode < 1000 || mode % 10 == 0 || mode % 10 == 3) && reps > 0 && cycles_dev,
                "umma_rate: bad arguments");
  probe::umma_rate_kernel<<<2, 128, 0, (cudaStream_t)stream>>>(mode, reps, cycles_dev);
  ISO_CHECK_LAUNCH("umma_rate_kernel");
  return ISOB200_OK;
}

extern "C" int isob200_umma2_probe(const float* a, const float* b, int K, float* dump, void* stream) {
  using namespace isob200;
  ISO_CHECK_ARG(a && b && dump, "umma2_probe: null pointer");
  ISO_CHECK_ARG(K == 16 || K == 32 || K == 48 || K == 64, "umma2_probe: K must be 16, 32, 48 or 64");
  probe::umma2_probe_kernel<<<2, 128, 0, (cudaStream_t)stream>>>(a, b, K, dump);
  ISO_CHECK_LAUNCH("umma2_probe_kernel");
  return ISOB200_OK;
}
