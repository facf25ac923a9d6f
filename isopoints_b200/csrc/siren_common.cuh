// Shared pieces of the fused SIREN kernel (siren.cu): packed-blob layout, PTX wrappers, packed fp32x2 math,
// fp16 hi/lo split, the fused-Newton argument block.
#pragma once
#include "common.cuh"
#include "isob200.h"
#include <cuda_fp16.h>

namespace isob200 {
namespace siren {

constexpr int H = 256;            // hidden width
constexpr int TM = 128;           // points per tile (UMMA M)
constexpr int KB = 32;            // k-block width (one weight stage)
constexpr int NKB = H / KB;       // 8 k-blocks per GEMM
constexpr int STAGES = 3;
constexpr int STAGE_PART = H * KB * 2;        // 16 KB: one fp16 part (hi or lo) of a k-block of B
constexpr int STAGE_BYTES = 2 * STAGE_PART;   // 32 KB: hi | lo
constexpr int A_PART = TM * H * 2;            // 64 KB: one fp16 part of the A tile
constexpr int A_LBO = TM * 16;                // 2048: byte stride between K-chunks (8 elems) of A
constexpr int B_LBO = H * 16;                 // 4096: same for a B stage
constexpr int SBO = 128;                      // byte stride between 8-row groups
constexpr int N_EPI_WARPS = 16;
constexpr int THREADS = (N_EPI_WARPS + 2) * 32;
constexpr float A_SCALE = 4096.f;             // 2^12: static scale of sin() activations
constexpr float A_SCALE_INV = 1.f / 4096.f;
constexpr int MAX_LAYERS = 32;

// ---- packed blob layout (all offsets in bytes) -------------------------------------------
// [0,1024)        header floats: [l-1] = 2^-s_l (inverse weight scale of hidden layer l), l = 1..L
//                 [32 + l-1] = |omega| * max_j sum_k |W_l[k][j]| (bound on the gain of the reverse GEMM of layer l:
//                 max_j |(gp W_l)_j| * |omega| <= max_k |gp_k| * this), l = 1..L
//                 [64] gl_scale, [65] 1/gl_scale (static scale of gp_L), [66] b_last,
//                 [67] omega_0 (first layer), [68] omega (hidden)
// [1024,5120)     float4 w0p[128][2]: column pair (a, b) = (2p, 2p+1) as (x_a, x_b, y_a, y_b), (z_a, z_b, w_a, w_b)
//                 with (x, y, z, w)_n = omega_0 * (W0[n,0], W0[n,1], W0[n,2], b0[n])
// [5120,6144)     float  w_last[256]
// [6144, ..)      float  bias[L][256], pre-multiplied by omega
// images          (1024-aligned) for l = 1..L, orientation o = 0 (forward: B[n][k] = W_l[n][k])
//                 and o = 1 (backward: B[n][k] = W_l[k][n]): 8 stages of 32 KB
constexpr size_t HDR_GAIN = 32;
constexpr size_t HDR_GL_SCALE = 64, HDR_GL_SCALE_INV = 65, HDR_B_LAST = 66, HDR_OMEGA0 = 67, HDR_OMEGA = 68;
constexpr size_t OFF_W0B = 1024;
constexpr size_t OFF_WLAST = OFF_W0B + H * 16;
constexpr size_t OFF_BIAS = OFF_WLAST + H * 4;
__host__ __device__ inline size_t off_images(int L) { return (OFF_BIAS + (size_t)L * H * 4 + 1023) / 1024 * 1024; }
__host__ __device__ inline size_t image_bytes() { return (size_t)NKB * STAGE_BYTES; }  // per (layer, orientation)
__host__ __device__ inline size_t blob_bytes(int L) { return off_images(L) + (size_t)L * 2 * image_bytes(); }
// scratch for the per-layer maxima (uint bit patterns of non-negative floats): L hidden + w_last
constexpr size_t PACK_WS_BYTES = (MAX_LAYERS + 1) * sizeof(unsigned);

// ---- shared memory map of the main kernel --------------------------------------------------
constexpr int SM_A_HI = 0;
constexpr int SM_A_LO = A_PART;
constexpr int SM_STAGE = 2 * A_PART;                        // 131072
constexpr int SM_XCH = SM_STAGE + STAGES * STAGE_BYTES;     // 229376: float xch[128][4]
constexpr int SM_BAR = SM_XCH + TM * 4 * 4;                 // 231424
// barriers: a_ready[8], acc_full[2], w_full[3], w_empty[3], tile[2]  -> 18 * 8 B
constexpr int SM_TMEM_PTR = SM_BAR + 18 * 8;
constexpr int SM_CMP = SM_TMEM_PTR + 16;                    // int cmp[8]: warp counts + reserved base
constexpr int SM_HDR = SM_CMP + 32;                         // float hdr[64]: per-layer scales / gains (header copy)
constexpr int SM_TILE = SM_HDR + 64 * 4;                    // int tile[2]: the scheduler's hand-off slots
constexpr int SMEM_BYTES = SM_TILE + 16;                    // 231888 <= 232448

// ---- PTX helpers -----------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// Bounded spin: a protocol bug traps (launch failure reported through the C ABI) instead of
// hanging the GPU.  2^24 polls is >= 0.3 s, far beyond any legitimate wait in this kernel.
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
    if (spin > (1u << 24)) __trap();
  }
}
__device__ __forceinline__ void tma_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// 32 lanes x 32 consecutive 32-bit columns: thread t of the warp receives lane (base lane + t)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
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
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
// 32 lanes x 8 consecutive columns; issue and wait are separate so the next k-block's load can be
// in flight during the current one's math.  The wait names the registers so that no use is
// scheduled above it.
__device__ __forceinline__ void tmem_ld8_issue(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_ld_wait(uint32_t (&r)[8]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7])
               :
               : "memory");
}
// barrier among the 4 epilogue warps that own the same 32 rows (one per 8-column slice)
__device__ __forceinline__ void row_barrier(int q) { asm volatile("bar.sync %0, 128;" ::"r"(q + 1) : "memory"); }

// K-major, no-swizzle shared-memory matrix descriptor (cute::UMMA::SmemDescriptor, version 1):
// core matrix = 8 rows x 16 B contiguous; SBO = stride between 8-row groups, LBO = stride
// between the two 16-byte K-chunks of one K=16 step.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(SBO >> 4) << 32) |
         (1ull << 46);
}
// instruction descriptor: D = f32, A = B = f16, K-major both, N = 256, M = 128
constexpr uint32_t IDESC = (1u << 4) | ((uint32_t)(H >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);

// ---- packed fp32x2 math (sm_100 FFMA2 / FMUL2 / FADD2: two lanes per issued instruction) ------
// The epilogue is instruction-issue bound (one row x 8 columns of sin/cos per thread and k-block),
// so every elementwise step works on column PAIRS.
__device__ __forceinline__ float2 bc2(float a) { return make_float2(a, a); }

// sin of two fp32 arguments, ~1 ulp absolute (1.2e-7 measured): j = round(a / pi) with the 1.5 * 2^23 trick,
// two-constant Cody-Waite reduction by pi (3.140625 has 9 significant bits, so j * c1 is exact for |j| < 2^15;
// the next term of pi, 5.1e-12 * j, stays below 2e-8 for |a| < 10^4 and is dropped), minimax polynomial on
// [-pi/2, pi/2] (degree 11 odd, least-squares fit on Chebyshev nodes).  Also returns the reduced argument r
// (a = j pi + r) and the parity of j as a sign bit (bit 31), from which the caller gets the cosine:
// cos(a) = cos(r) ^ sg (cos_of_reduced), or stores both in one word for the reverse pass (tape_word).
__device__ __forceinline__ void sin_red2(float2 a, float2& sn, float2& r, uint32_t& sgx, uint32_t& sgy) {
  const float2 jm = __ffma2_rn(a, bc2(0.318309886f), bc2(12582912.f));
  sgx = __float_as_uint(jm.x) << 31;
  sgy = __float_as_uint(jm.y) << 31;
  const float2 j = __fadd2_rn(jm, bc2(-12582912.f));
  r = __ffma2_rn(j, bc2(-3.140625f), a);
  r = __ffma2_rn(j, bc2(-9.676535846665502e-4f), r);
  const float2 r2 = __fmul2_rn(r, r);
  float2 t = __ffma2_rn(r2, bc2(-2.39068338458992e-08f), bc2(2.7526464236871107e-06f));
  t = __ffma2_rn(t, r2, bc2(-1.9840890308842063e-04f));
  t = __ffma2_rn(t, r2, bc2(8.333330973982811e-03f));
  t = __ffma2_rn(t, r2, bc2(-0.1666666716337204f));
  t = __fmul2_rn(t, r2);
  const float2 rs = make_float2(__uint_as_float(__float_as_uint(r.x) ^ sgx), __uint_as_float(__float_as_uint(r.y) ^ sgy));
  sn = __ffma2_rn(t, rs, rs);
}
// reduced argument and parity sign of a alone (no sine): for the stages that only need the cosine
__device__ __forceinline__ void red2(float2 a, float2& r, uint32_t& sgx, uint32_t& sgy) {
  const float2 jm = __ffma2_rn(a, bc2(0.318309886f), bc2(12582912.f));
  sgx = __float_as_uint(jm.x) << 31;
  sgy = __float_as_uint(jm.y) << 31;
  const float2 j = __fadd2_rn(jm, bc2(-12582912.f));
  r = __ffma2_rn(j, bc2(-3.140625f), a);
  r = __ffma2_rn(j, bc2(-9.676535846665502e-4f), r);
}
// |cos| branch of the reduced argument (|r| <= pi/2, so the value is >= 0 up to rounding of r): MUFU.COS, absolute
// error <= 2^-21.2 there.  The cosine only ever multiplies back-propagated rows (the tape factor w cos(w z)); the
// forward value path uses the sine alone.  The epilogue is issue / FP32-pipe bound (a packed FFMA2 holds the pipe for
// two cycles), so the cosine runs on the special-function unit instead of a six-term polynomial.
__device__ __forceinline__ float2 cos_of_reduced(float2 r) { return make_float2(__cosf(r.x), __cosf(r.y)); }
// Reverse-mode tape entry: |r| with the parity of j in the sign bit (one LOP3).  cos is even, so the reverse pass
// takes MUFU.COS of the word as it is and xors the word's sign bit into the result: cos(a) = cos(|r|) ^ parity.
// The cosine itself is NOT evaluated in the forward pass -- the forward epilogue paces its GEMMs, the reverse one
// waits for them.
__device__ __forceinline__ float tape_word(float r, uint32_t sg) {
  return __uint_as_float((__float_as_uint(r) & 0x7fffffffu) | sg);
}
__device__ __forceinline__ float2 cos_of_tape(float wx, float wy) {
  const float cx = __cosf(wx), cy = __cosf(wy);
  return make_float2(__uint_as_float(__float_as_uint(cx) ^ (__float_as_uint(wx) & 0x80000000u)),
                     __uint_as_float(__float_as_uint(cy) ^ (__float_as_uint(wy) & 0x80000000u)));
}
__device__ __forceinline__ float2 signed_scale(float sc, uint32_t sgx, uint32_t sgy) {
  return make_float2(__uint_as_float(__float_as_uint(sc) ^ sgx), __uint_as_float(__float_as_uint(sc) ^ sgy));
}

// split 8 scaled fp32 values (4 pairs) into fp16 hi / lo and store the two 16-byte K-chunks
__device__ __forceinline__ void store_chunk(uint32_t a_hi_addr, uint32_t a_lo_addr, const float2* o) {
  uint32_t hi[4], lo[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    __half2 h = __floats2half2_rn(o[i].x, o[i].y);
    const float2 d = __ffma2_rn(__half22float2(h), bc2(-1.f), o[i]);
    __half2 l = __floats2half2_rn(d.x, d.y);
    hi[i] = *reinterpret_cast<uint32_t*>(&h);
    lo[i] = *reinterpret_cast<uint32_t*>(&l);
  }
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a_hi_addr), "r"(hi[0]), "r"(hi[1]), "r"(hi[2]),
               "r"(hi[3])
               : "memory");
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a_lo_addr), "r"(lo[0]), "r"(lo[1]), "r"(lo[2]),
               "r"(lo[3])
               : "memory");
}

// power of two p with bound * p in [2^11, 2^12) (1 for bound == 0 / non-finite)
__device__ __forceinline__ float pow2_scale_for(float bound) {
  int e = (int)((__float_as_uint(bound) >> 23) & 0xFF);  // biased exponent
  if (e == 0 || e == 255) return 1.f;
  int se = 127 + 11 - (e - 127);  // biased exponent of 2^(11 - (e-127))
  se = se < 1 ? 1 : (se > 254 ? 254 : se);
  return __uint_as_float((uint32_t)se << 23);
}


// Optional fused Newton step (UniformProjection._project_points, levelset_sampling.py:313-342): when
// `points` is non-null the thread that ends up holding a row's sdf and gradient also applies the update
// of csrc/project.cu's project_step_kernel (same fp32 operation order) and appends still-active rows to
// the next iteration's active list -- the SDF value and gradient never leave the SM.
struct Newton {
  float* points;             // (M,3) packed positions, updated in place at the active rows
  float* normals;            // (M,3) last gradient
  unsigned char* not_conv;   // (M) flags
  const int* act_in;         // active row ids of this launch (NULL = identity)
  int* act_out;              // still-active row ids (tiles append in completion order)
  float* next_points;        // their updated positions, same order (NULL on the last evaluation)
  int* count_out;            // number of still-active rows (zero before the launch)
  float tol, max_step;
  int do_update;
  // mode 1: ray marching (SphereTracing.project_points, levelset_sampling.py:733-786) on the forward half of
  // the network only -- no gradient, no tape: rays advance by alpha * sdf along dirs, tol = 0.1 * proj_tolerance
  int mode;
  const float* dirs;         // (M,3) ray directions
  float* eval;               // (M) last sdf per ray
  float alpha, bound;        // step factor; radius + padding of the bounding sphere
  // tile scheduler: slot of g_tile_sched used by this launch; CTAs on odd SMs start stagger_cycles late when
  // the launch has at least stagger_min_tiles tiles (see the producer warp)
  int sched_slot, stagger_cycles, stagger_min_tiles;
};

// eps_denom(x, eps) of DSS/utils/mathHelper.py:14-18: (sign(x) + [x == 0]) * max(|x|, eps)
__device__ __forceinline__ float eps_denom_f(float x, float eps) {
  const float sgn = (x > 0.f) ? 1.f : ((x < 0.f) ? -1.f : 1.f);
  return __fmul_rn(sgn, fmaxf(fabsf(x), eps));
}

// MMA k-block visiting order = the order in which the epilogue completes them.

}  // namespace siren
}  // namespace isob200
