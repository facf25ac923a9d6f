// Fused SIREN SDF value + input-gradient kernel.
//
// Replaces, for SDF modules that are a plain SIREN MLP (DSS/models/common.py:56-165: SineLayer
// chain + linear head), the autograd round trip of UniformProjection._compute_sdf_and_grad
// (DSS/models/levelset_sampling.py:142-170): model.forward(x).sdf followed by
// autograd.grad(sdf, x, ones).  One persistent kernel evaluates the whole network and its
// reverse-mode input gradient for 128-point tiles:
//
//   forward   z_0 = x W_0^T + b_0 (K = 3, SIMT)            h_0 = sin(w0 z_0)
//             z_l = h_{l-1} W_l^T + b_l  (tcgen05 GEMM)    h_l = sin(w z_l),  c_l = w cos(w z_l) (taken in the backward pass)
//             sdf = h_L . w_last + b_last
//   backward  gp_L = w_last * c_L ;  g_{l-1} = gp_l W_l (tcgen05 GEMM) ;  gp_{l-1} = g_{l-1} * c_{l-1}
//             grad = gp_0 W_0
//
// B200 mapping
//   * the 2L hidden GEMMs [128 x 256] x [256 x 256] run on the 5th-gen tensor cores
//     (tcgen05.mma.cta_group::1.kind::f16, M=128, N=256, K=16), accumulators in TMEM (two
//     256-column buffers, ping-pong between consecutive GEMMs);
//   * fp32 accuracy from fp16 tensor-core products: every operand is pre-scaled by a power of
//     two to ~2^12 and split hi + lo (two fp16, 22 significand bits); each k-step issues
//     hi*hi + lo*hi + hi*lo into the same fp32 accumulator (the dropped lo*lo term is 2^-24
//     relative).  Activations are scaled by the static 2^12, weights per layer by a power of
//     two derived on the device from max|W|, back-propagated rows by a per-row power of two;
//     all scalings are exact and undone in the epilogue;
//   * the A operand never leaves the SM: the epilogue warps read the accumulator with
//     tcgen05.ld, apply bias / sin / cos (or the cos factor in the backward pass), split to
//     fp16 hi/lo and store straight into the next GEMM's canonical K-major shared-memory
//     layout (no-swizzle core matrices).  The hand-off is per 32-column k-block (8 mbarriers),
//     so the next GEMM's tensor-core work overlaps the rest of the epilogue;
//   * weights are packed once per parameter version (siren_pack_kernel) into the exact
//     shared-memory image of each 32-wide k-block stage (hi | lo, 32 KB) for both orientations
//     (W for the forward pass, W^T for the backward pass) and streamed L2 -> smem with 1-D TMA
//     bulk copies (cp.async.bulk + mbarrier complete_tx) through a 3-stage ring by a producer
//     warp;
//   * the reverse-mode tape is NOT the cosine: the forward epilogue (which paces its GEMM) stores the
//     reduced argument of each sine with the parity of the reduction in its sign bit, and the reverse
//     epilogue (which waits for its GEMM) takes MUFU.COS of that word.  The tape lives in a per-CTA
//     scratch slab in global memory (float4 per thread, 512 B contiguous per warp): written and read
//     back by the same thread within ~100 us, prefetched into L2 one stage ahead, discarded from L2
//     after its last use;
//   * tiles are handed out dynamically (one atomic per tile on a self-resetting device table).
//   Warp roles: warps 0-15 epilogue (TMEM lane quarter = warp % 4; within every 32-column k-block
//   the warp owns the 8-column slice warp / 4, i.e. exactly one 16-byte K-chunk of its row, so the
//   k-blocks of the next GEMM's A operand complete one after the other and the tensor core trails
//   the epilogue by a single k-block), warp 16 tile scheduler + weight producer, warp 17 TMEM
//   allocator + single-thread MMA issuer.
//   What bounds it in a sustained run is the board's power limit, not cycles.
//
// Only H = 256 is built (BASELINE C2's "8-layer x 256 SIREN"); other widths keep the autograd path.
#include "siren_common.cuh"
#include <atomic>

namespace isob200 {
namespace siren {

// ---------------------------------------------------------------------------------------------
// pack: per-layer maxima, then fp16 hi/lo stage images + the small fp32 tables
// ---------------------------------------------------------------------------------------------
__global__ void siren_absmax_kernel(const float* __restrict__ w_hidden, const float* __restrict__ w_last, int L,
                                    unsigned* __restrict__ maxbits) {
  // blockIdx.y = layer (L = w_last)
  int l = blockIdx.y;
  const float* src = l < L ? w_hidden + (size_t)l * H * H : w_last;
  int n = l < L ? H * H : H;
  float m = 0.f;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) m = fmaxf(m, fabsf(src[i]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(maxbits + l, __float_as_uint(m));
}

// hdr[HDR_GAIN + l] = |omega| * max_j sum_k |W_l[k][j]|  (one block per hidden layer, one thread per column j)
__global__ void siren_gain_kernel(const float* __restrict__ w_hidden, float omega, unsigned char* __restrict__ blob) {
  const int l = blockIdx.x, j = threadIdx.x;
  const float* W = w_hidden + (size_t)l * H * H;
  float s = 0.f;
  for (int k = 0; k < H; ++k) s += fabsf(W[(size_t)k * H + j]);
  __shared__ float red[H / 32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s = fmaxf(s, __shfl_xor_sync(0xffffffffu, s, o));
  if ((j & 31) == 0) red[j >> 5] = s;
  __syncthreads();
  if (j == 0) {
    float m = red[0];
    for (int i = 1; i < H / 32; ++i) m = fmaxf(m, red[i]);
    reinterpret_cast<float*>(blob)[HDR_GAIN + l] = m * fabsf(omega);
  }
}

__global__ void siren_pack_kernel(const float* __restrict__ w0, const float* __restrict__ b0,
                                  const float* __restrict__ w_hidden, const float* __restrict__ b_hidden,
                                  const float* __restrict__ w_last, const float* __restrict__ b_last, float omega0,
                                  float omega, int L, const unsigned* __restrict__ maxbits,
                                  unsigned char* __restrict__ blob) {
  float* hdr = reinterpret_cast<float*>(blob);
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  const int nth = gridDim.x * blockDim.x;
  if (tid < L) hdr[tid] = 1.f / pow2_scale_for(__uint_as_float(maxbits[tid]));
  if (tid == 0) {
    float gl = pow2_scale_for(fabsf(omega) * __uint_as_float(maxbits[L]));
    hdr[HDR_GL_SCALE] = gl;
    hdr[HDR_GL_SCALE_INV] = 1.f / gl;
    hdr[HDR_B_LAST] = b_last ? b_last[0] : 0.f;
    hdr[HDR_OMEGA0] = omega0;
    hdr[HDR_OMEGA] = omega;
  }
  float4* w0b = reinterpret_cast<float4*>(blob + OFF_W0B);
  float* wl = reinterpret_cast<float*>(blob + OFF_WLAST);
  float* bias = reinterpret_cast<float*>(blob + OFF_BIAS);
  for (int i = tid; i < H; i += nth) {
    {
      float* w0p = reinterpret_cast<float*>(w0b) + (i >> 1) * 8 + (i & 1);   // pair-transposed, see layout
      w0p[0] = omega0 * w0[3 * i];
      w0p[2] = omega0 * w0[3 * i + 1];
      w0p[4] = omega0 * w0[3 * i + 2];
      w0p[6] = b0 ? omega0 * b0[i] : 0.f;
    }
    wl[i] = w_last[i];
  }
  for (int i = tid; i < L * H; i += nth) bias[i] = b_hidden ? omega * b_hidden[i] : 0.f;
  // images: one thread per 16-byte chunk (8 consecutive k of one row n)
  const size_t img0 = off_images(L);
  const int chunks_per_img = NKB * (KB / 8) * H;  // 8 * 4 * 256
  const long long total = (long long)L * 2 * chunks_per_img;
  for (long long c = tid; c < total; c += nth) {
    int n = (int)(c % H);
    int k8 = (int)((c / H) % (KB / 8));
    int kb = (int)((c / (H * (KB / 8))) % NKB);
    int o = (int)((c / chunks_per_img) % 2);
    int l = (int)(c / (2 * chunks_per_img));
    const float* W = w_hidden + (size_t)l * H * H;
    float sc = pow2_scale_for(__uint_as_float(maxbits[l]));
    int k0 = kb * KB + k8 * 8;
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float v0 = (o == 0 ? W[(size_t)n * H + k0 + 2 * i] : W[(size_t)(k0 + 2 * i) * H + n]) * sc;
      float v1 = (o == 0 ? W[(size_t)n * H + k0 + 2 * i + 1] : W[(size_t)(k0 + 2 * i + 1) * H + n]) * sc;
      __half2 h = __floats2half2_rn(v0, v1);
      float2 f = __half22float2(h);
      __half2 lw = __floats2half2_rn(v0 - f.x, v1 - f.y);
      hi[i] = *reinterpret_cast<uint32_t*>(&h);
      lo[i] = *reinterpret_cast<uint32_t*>(&lw);
    }
    unsigned char* stage = blob + img0 + ((size_t)(l * 2 + o) * NKB + kb) * STAGE_BYTES;
    size_t off = (size_t)k8 * B_LBO + (size_t)(n >> 3) * SBO + (size_t)(n & 7) * 16;
    *reinterpret_cast<uint4*>(stage + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(stage + STAGE_PART + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  }
}

// ld.global.cg issued exactly here (volatile: ptxas does not sink it to the first use, which would put the L2
// round trip back on the path it is meant to be taken off)
__device__ __forceinline__ float4 ldcg_now(const float4* p) {
  float4 v;
  asm volatile("ld.global.cg.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ float4 ldnc_now(const float4* p) {
  float4 v;
  asm volatile("ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
// ---------------------------------------------------------------------------------------------
// main kernel
// ---------------------------------------------------------------------------------------------
// Tile scheduler state: [slot][0] = next tile, [slot][1] = CTAs finished.  Zero at module load; the last CTA of a
// launch puts its slot back to zero, and consecutive launches take consecutive slots, so kernels that overlap on
// different streams do not share one.
constexpr int SCHED_SLOTS = 64;
__device__ int g_tile_sched[SCHED_SLOTS * 2];

#define ISO_STAMP(p, i)              \
  if constexpr (DBG) {               \
    if (p) (p)[i] = clock64();       \
  }

// DBG = true: the bring-up build of the same kernel (cycle stamps of CTA 0's first two tiles for dbg_gemm == -2, raw
// accumulator dump of GEMM dbg_gemm >= 0); launched only when the caller passes a dbg buffer.  The product
// instantiation carries none of those instructions (they were ~6 % of the epilogue's issue slots when predicated off).
template <bool DBG>
__global__ void __launch_bounds__(THREADS, 1)
siren_sdf_grad_kernel(const float* __restrict__ x, int n_max, const int* __restrict__ n_dev,
                      const unsigned char* __restrict__ blob, int L, float* __restrict__ sdf_out,
                      float* __restrict__ grad_out, float* __restrict__ scratch, float* __restrict__ dbg,
                      int dbg_gemm, const Newton nw) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bar_a_ready = sbase + SM_BAR;             // [8]
  const uint32_t bar_acc_full = sbase + SM_BAR + 8 * 8;    // [2]
  const uint32_t bar_w_full = sbase + SM_BAR + 10 * 8;     // [3]
  const uint32_t bar_w_empty = sbase + SM_BAR + 13 * 8;    // [3]
  const uint32_t bar_tile = sbase + SM_BAR + 16 * 8;       // [2]
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(smem + SM_TMEM_PTR);
  volatile int* tile_slot = reinterpret_cast<volatile int*>(smem + SM_TILE);   // [2]

  int n = n_max;
  if (n_dev) {
    int nd = *n_dev;
    n = nd < n_max ? nd : n_max;
  }
  const int num_tiles = (n + TM - 1) / TM;

  if (threadIdx.x == 0) {
    for (int i = 0; i < NKB; ++i) mbar_init(bar_a_ready + 8 * i, N_EPI_WARPS);
    for (int i = 0; i < 2; ++i) mbar_init(bar_acc_full + 8 * i, 1);
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(bar_w_full + 8 * i, 1);
      mbar_init(bar_w_empty + 8 * i, 1);
    }
    for (int i = 0; i < 2; ++i) mbar_init(bar_tile + 8 * i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 64)   // per-layer scale / gain table: read at the start of every stage, so keep it one LDS away
    reinterpret_cast<float*>(smem + SM_HDR)[threadIdx.x] = reinterpret_cast<const float*>(blob)[threadIdx.x];
  if (warp == N_EPI_WARPS + 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sbase + SM_TMEM_PTR),
                 "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  const float* hdr = reinterpret_cast<const float*>(blob);
  const size_t img0 = off_images(L);
  const bool fwd_only = nw.mode != 0;            // value only (mode 2) / ray marching (mode 1): L GEMMs, no tape
  const int n_gemm = fwd_only ? L : 2 * L;      // per tile
  int* sched = g_tile_sched + 2 * nw.sched_slot;

  if (warp == N_EPI_WARPS) {
    // ===================== tile scheduler + weight producer =====================
    if (lane == 0) {
      // Phase stagger.  A tile's reverse-mode tape grows to (L-1) x 128 KB over its forward half and is consumed over
      // its reverse half; with every CTA in the same phase the live tape peaks at 148 x 768 KB = 116 MB (L = 7), of
      // which only ~60 MB stay in L2 -- the rest makes a round trip through HBM (0.95 GB per 200 k rows, measured)
      // and the reverse stages wait for it.  Half of the CTAs (odd SM id: the two SMs of a TPC differ) therefore
      // start half a tile period late: one half of the chip is writing tape while the other half consumes it, the
      // live total stays at its 55 MB mean and the kernel time per wave drops by ~12 % (89 vs 101 us, the time a
      // 74-CTA launch needs per wave).  Tiles are handed out dynamically, so a late CTA simply takes fewer of them.
      if (!fwd_only && nw.stagger_cycles > 0 && num_tiles >= nw.stagger_min_tiles) {
        unsigned smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        if (smid & 1u) {
          const long long t0 = clock64();
          while (clock64() - t0 < (long long)nw.stagger_cycles) __nanosleep(512);
        }
      }
      uint32_t it = 0;
      for (uint32_t k = 0;; ++k) {
        int tile = atomicAdd(sched, 1);
        if (tile >= num_tiles) tile = -1;
        tile_slot[k & 1] = tile;
        mbar_arrive(bar_tile + 8 * (k & 1));   // release: the slot is visible to whoever sees the phase complete
        if (tile < 0) break;
        for (int g = 0; g < n_gemm; ++g) {
          // forward GEMM g -> layer g+1, orientation 0 ; backward -> layer 2L-g, orientation 1
          int l = g < L ? g : (2 * L - 1 - g);  // 0-based hidden layer index
          int o = g < L ? 0 : 1;
          const unsigned char* img = blob + img0 + (size_t)(l * 2 + o) * image_bytes();
          for (int i = 0; i < NKB; ++i, ++it) {
            uint32_t s = it % STAGES;
            uint32_t ph = (it / STAGES) & 1;
            mbar_wait(bar_w_empty + 8 * s, ph ^ 1);
            if constexpr (DBG) {
              // dbg_gemm == -3 (energy experiment, wrong results): the weight stream is skipped after the first
              // pass over the ring -- what the 4.7 TB/s of L2 -> shared-memory traffic costs under the power cap
              if (dbg_gemm == -3 && it >= STAGES) {
                mbar_arrive(bar_w_full + 8 * s);
                continue;
              }
            }
            mbar_expect_tx(bar_w_full + 8 * s, STAGE_BYTES);
            tma_bulk_g2s(sbase + SM_STAGE + s * STAGE_BYTES, img + (size_t)i * STAGE_BYTES, STAGE_BYTES,
                         bar_w_full + 8 * s);
          }
          // The tape words the reverse stage after next will read were written up to 2(L-1) stages ago and may
          // have left L2: ask the TMA unit to pull that layer's whole 128 KB slab ([col4][row] float4, contiguous)
          // back in -- one instruction from this otherwise idle thread instead of 16 prefetches per epilogue thread.
          // (GEMM g's weights are all requested: the epilogue is about to start stage g - 1 or g.)
          if (!fwd_only) {
            const int lt = g < L ? (g + 1 == L ? L - 1 : 0) : (2 * L - g) - 2;   // 1-based layer of that slab
            if (lt >= 1)
              asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(
                               reinterpret_cast<const float4*>(scratch) +
                               ((size_t)blockIdx.x * (size_t)(L > 1 ? L - 1 : 1) + (size_t)(lt - 1)) * 64 * TM),
                           "r"(64 * TM * 16)
                           : "memory");
          }
        }
      }
    }
  } else if (warp == N_EPI_WARPS + 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      uint32_t it = 0;
      uint32_t G = 0;  // global GEMM counter of this CTA
      // dbg_gemm == -2: cycle stamps of the first two tiles of CTA 0 (bring-up / tuning aid)
      long long* tstamp = nullptr;
      if constexpr (DBG) tstamp = (dbg && dbg_gemm == -2 && blockIdx.x == 0) ? reinterpret_cast<long long*>(dbg) : nullptr;
      for (uint32_t k = 0;; ++k) {
        mbar_wait(bar_tile + 8 * (k & 1), (k >> 1) & 1);
        if (tile_slot[k & 1] < 0) break;
        for (int g = 0; g < n_gemm; ++g, ++G) {
          const uint32_t d_tmem = tmem_base + (G & 1) * H;
          long long wa = 0, ww = 0, t_first = 0;
          for (int i = 0; i < NKB; ++i, ++it) {
            uint32_t s = it % STAGES;
            uint32_t ph = (it / STAGES) & 1;
            // weights first: they are normally in place already, and a completed try_wait still costs ~180
            // cycles -- this keeps it off the path between the epilogue's hand-off and the MMA issue
            long long t0 = 0, t1 = 0;
            if constexpr (DBG) t0 = tstamp ? clock64() : 0;
            mbar_wait(bar_w_full + 8 * s, ph);
            if constexpr (DBG) t1 = tstamp ? clock64() : 0;
            mbar_wait(bar_a_ready + 8 * i, G & 1);
            if constexpr (DBG) {
              if (tstamp) {
                long long t2 = clock64();
                ww += t1 - t0;
                wa += t2 - t1;
                if (i == 0) t_first = t2;
              }
            }
            tc_fence_after();
            const uint32_t a_hi = sbase + SM_A_HI + i * (KB / 8) * A_LBO;
            const uint32_t a_lo = sbase + SM_A_LO + i * (KB / 8) * A_LBO;
            const uint32_t b_hi = sbase + SM_STAGE + s * STAGE_BYTES;
            const uint32_t b_lo = b_hi + STAGE_PART;
#pragma unroll
            for (int k16 = 0; k16 < KB / 16; ++k16) {
              uint64_t dah = make_desc(a_hi + k16 * 2 * A_LBO, A_LBO);
              uint64_t dal = make_desc(a_lo + k16 * 2 * A_LBO, A_LBO);
              uint64_t dbh = make_desc(b_hi + k16 * 2 * B_LBO, B_LBO);
              uint64_t dbl = make_desc(b_lo + k16 * 2 * B_LBO, B_LBO);
              bool one_product = false;   // dbg_gemm == -6 (energy experiment, wrong results): hi*hi only
              if constexpr (DBG) one_product = dbg_gemm == -6;
              if (!one_product) {
                tc_mma_f16(d_tmem, dal, dbh, IDESC, (i | k16) ? 1u : 0u);
                tc_mma_f16(d_tmem, dah, dbl, IDESC, 1u);
              }
              tc_mma_f16(d_tmem, dah, dbh, IDESC, (one_product && !(i | k16)) ? 0u : 1u);
            }
            tc_commit(bar_w_empty + 8 * s);  // stage reusable once these MMAs have read it
          }
          tc_commit(bar_acc_full + 8 * (G & 1));  // accumulator of GEMM G complete
          if constexpr (DBG) {
            if (tstamp && G < 2u * n_gemm) {
              tstamp[G * 8 + 3] = t_first;
              tstamp[G * 8 + 4] = clock64();
              tstamp[G * 8 + 5] = ww;
              tstamp[G * 8 + 6] = wa;
            }
          }
        }
      }
    }
  } else {
    // ===================== epilogue warps =====================
    const int q = warp & 3;        // TMEM lane quarter
    const int cslice = warp >> 2;  // 8-column slice inside every 32-column k-block
    const int row = 32 * q + lane;
    const uint32_t tl = tmem_base + ((uint32_t)(32 * q) << 16) + 8 * cslice;
    // this thread's 16-byte K-chunk of k-block kb lives at chunk index 4 kb + cslice
    const uint32_t a_thr = (uint32_t)cslice * A_LBO + (uint32_t)(row >> 3) * SBO + (uint32_t)(row & 7) * 16;
    float* xch = reinterpret_cast<float*>(smem + SM_XCH);
    const float* shdr = reinterpret_cast<const float*>(smem + SM_HDR);
    // grad partials of column slices 1..3 go through the (idle) A tile in the last stage: the 16-byte
    // slot of this row in k-chunk (cslice - 1), which only this row's own warps ever write
    unsigned char* gsc = smem + SM_A_HI + row * 16;
    const float4* w0p = reinterpret_cast<const float4*>(blob + OFF_W0B) + cslice * 8;   // 4 pairs x 2 float4 per k-block
    const float4* w_last4 = reinterpret_cast<const float4*>(blob + OFF_WLAST) + cslice * 2;
    const float* biasw = reinterpret_cast<const float*>(blob + OFF_BIAS);
    const float omega = hdr[HDR_OMEGA];
    const float gl_scale = hdr[HDR_GL_SCALE], gl_scale_inv = hdr[HDR_GL_SCALE_INV];
    const float b_last = hdr[HDR_B_LAST];
    // per-CTA stash: [(l-1)][col4 (64)][row (128)] float4, l = 1..L-1
    float4* stash = reinterpret_cast<float4*>(scratch) + (size_t)blockIdx.x * (size_t)(L > 1 ? L - 1 : 1) * 64 * TM;
    uint32_t G = 0;

    // publish this thread's K-chunk of k-block kb: generic-proxy stores -> async proxy, one arrive per warp
    auto publish = [&](int kb, const float2* o, long long* kst = nullptr) {
      const uint32_t off = (uint32_t)(kb * 4) * A_LBO + a_thr;
      store_chunk(sbase + SM_A_HI + off, sbase + SM_A_LO + off, o);
      ISO_STAMP(kst, 2);
      tc_fence_before();
      fence_proxy_async();
      ISO_STAMP(kst, 3);
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_a_ready + 8 * kb);
      ISO_STAMP(kst, 4);
    };

    // append the still-active rows of this tile to the next active list: in-tile order, one reservation per
    // tile.  Called by the four cslice == 0 warps (they hold the tile's 128 rows), all lanes.
    auto append_active = [&](bool still, int p, float nx, float ny, float nz) {
      int* cmp = reinterpret_cast<int*>(smem + SM_CMP);
      const unsigned bal = __ballot_sync(0xffffffffu, still);
      if (lane == 0) cmp[q] = __popc(bal);
      asm volatile("bar.sync 5, 128;" ::: "memory");
      const int c0 = cmp[0], c1 = cmp[1], c2 = cmp[2], c3 = cmp[3];
      if (threadIdx.x == 0) cmp[4] = (c0 + c1 + c2 + c3) ? atomicAdd(nw.count_out, c0 + c1 + c2 + c3) : 0;
      asm volatile("bar.sync 5, 128;" ::: "memory");
      if (still) {
        const int pos = cmp[4] + (q > 0 ? c0 : 0) + (q > 1 ? c1 : 0) + (q > 2 ? c2 : 0) +
                        __popc(bal & ((1u << lane) - 1u));
        nw.act_out[pos] = p;
        if (nw.next_points) {
          nw.next_points[3 * (size_t)pos] = nx;
          nw.next_points[3 * (size_t)pos + 1] = ny;
          nw.next_points[3 * (size_t)pos + 2] = nz;
        }
      }
    };

    for (uint32_t kt = 0;; ++kt) {
      mbar_wait(bar_tile + 8 * (kt & 1), (kt >> 1) & 1);
      const int tile = tile_slot[kt & 1];
      if (tile < 0) break;
      const int grow = tile * TM + row;
      float px = 0.f, py = 0.f, pz = 0.f;
      if (grow < n) {
        px = x[3 * (size_t)grow];
        py = x[3 * (size_t)grow + 1];
        pz = x[3 * (size_t)grow + 2];
      }
      const float2 px2 = bc2(px), py2 = bc2(py), pz2 = bc2(pz);
      // ---- E0: first layer in SIMT, A = 2^12 sin(w0 z_0) ----
#pragma unroll 1
      for (int kb = 0; kb < NKB; ++kb) {
        float2 o[4];
#pragma unroll
        for (int pr = 0; pr < 4; ++pr) {
          const float4 wa = __ldg(w0p + kb * 32 + pr * 2), wb = __ldg(w0p + kb * 32 + pr * 2 + 1);
          const float2 th = __ffma2_rn(make_float2(wb.x, wb.y), pz2,
                                       __ffma2_rn(make_float2(wa.z, wa.w), py2,
                                                  __ffma2_rn(make_float2(wa.x, wa.y), px2, make_float2(wb.z, wb.w))));
          float2 sn, r;
          uint32_t sx, sy;
          sin_red2(th, sn, r, sx, sy);
          o[pr] = __fmul2_rn(sn, bc2(A_SCALE));
        }
        publish(kb, o);
      }

      // operands of the NEXT stage's first k-blocks, requested one stage ahead so that their L2 latency is not
      // between an accumulator becoming ready and the first k-block of the next A operand:
      float4 pf0, pf1, pf2, pf3;   // forward: biases of k-block 0 (pf0, pf1); reverse: tape words of k-blocks 0, 1
      pf0 = pf1 = pf2 = pf3 = make_float4(0.f, 0.f, 0.f, 0.f);
      if (L > 0) {
        const float4* bw4n = reinterpret_cast<const float4*>(biasw) + cslice * 2;   // layer 1
        pf0 = __ldg(bw4n);
        pf1 = __ldg(bw4n + 1);
      }
      float row_scale_inv = 1.f;  // inverse of the scale applied to this row of the current backward A
      float row_max = 0.f;        // max |.| over this row of the current backward A (scaled values, all 256 columns)
      float sdf_row = 0.f;        // this row's sdf (slice-0 thread), kept for the fused Newton step
      for (int g = 0; g < n_gemm; ++g, ++G) {
        const uint32_t buf = G & 1;
        long long* tstamp = nullptr;   // stage stamps (thread 0)
        long long* kbase = nullptr;    // per-k-block stamps of thread 0 in stages G = 1 (forward) and G = 8 (reverse)
        long long* wst = nullptr;      // per-warp stamps (lane 0) in the same two stages
        if constexpr (DBG) {
          const bool on = dbg && dbg_gemm == -2 && blockIdx.x == 0 && G < 2u * n_gemm;
          long long* d64 = reinterpret_cast<long long*>(dbg);
          if (on && threadIdx.x == 0) tstamp = d64 + G * 8;
          if (tstamp && (G == 1u || G == 8u)) kbase = d64 + 512 + (G == 1u ? 0 : 64);
          if (on && lane == 0 && (G == 1u || G == 8u)) wst = d64 + 1024 + (G == 1u ? 0 : 64) + warp * 4;
        }
        ISO_STAMP(tstamp, 7);
        mbar_wait(bar_acc_full + 8 * buf, (G >> 1) & 1);
        tc_fence_after();
        ISO_STAMP(tstamp, 0);
        ISO_STAMP(wst, 0);
        const uint32_t tacc = tl + buf * H;
        const bool fwd = g < L;
        const int l = fwd ? g + 1 : 2 * L - g;  // 1-based hidden layer this GEMM belongs to
        const float wsi = shdr[l - 1];
        // accumulator slices of consecutive k-blocks alternate between two register sets, each loaded one k-block
        // ahead of its use (the loops below are unrolled in pairs: no register-to-register copies)
        uint32_t rn[8], rm[8];
        tmem_ld8_issue(tacc, rn);
        if constexpr (DBG) {
          if (dbg && blockIdx.x == 0 && kt == 0 && (int)G == dbg_gemm) {
            // raw accumulator dump (unscaled), [128][256]
            tmem_ld_wait(rn);
#pragma unroll 1
            for (int kb = 0; kb < NKB; ++kb) {
              uint32_t r[8];
              tmem_ld8_issue(tacc + kb * KB, r);
              tmem_ld_wait(r);
#pragma unroll
              for (int i = 0; i < 8; ++i) dbg[(size_t)row * H + kb * KB + cslice * 8 + i] = __uint_as_float(r[i]);
            }
            tmem_ld8_issue(tacc, rn);
          }
        }

        if (fwd && l < L) {
          // ---- E_f(l): h_l = sin(w z_l) -> A ; tape_l = reduced argument + parity of w z_l ----
          const float2 sc2 = bc2(wsi * A_SCALE_INV * omega);
          float4* st = stash + (size_t)(l - 1) * 64 * TM + row + (size_t)(cslice * 2) * TM;
          const float4* bw4 = reinterpret_cast<const float4*>(biasw + (l - 1) * H) + cslice * 2;
          // one k-block: cur = this k-block's accumulator slice (in flight), nxt = the other register set;
          // (b0, b1) = this k-block's biases, (nb0, nb1) receive the next k-block's (layer l + 1's first at the end)
          auto fwd_block = [&](const int kb, uint32_t(&cur)[8], uint32_t(&nxt)[8], const float4& b0, const float4& b1,
                               float4& nb0, float4& nb1) {
            long long* kst = nullptr;
            if constexpr (DBG) kst = kbase ? kbase + kb * 8 : nullptr;
            ISO_STAMP(kst, 0);
            tmem_ld_wait(cur);
            ISO_STAMP(kst, 1);
            if (kb + 1 < NKB) tmem_ld8_issue(tacc + (kb + 1) * KB, nxt);
            const float2 bb[4] = {make_float2(b0.x, b0.y), make_float2(b0.z, b0.w), make_float2(b1.x, b1.y),
                                  make_float2(b1.z, b1.w)};
            float2 o[4];
            float tw[8];
#pragma unroll
            for (int pr = 0; pr < 4; ++pr) {
              const float2 v = make_float2(__uint_as_float(cur[2 * pr]), __uint_as_float(cur[2 * pr + 1]));
              float2 sn, r;
              uint32_t sx, sy;
              bool sfu_sine = false;   // dbg_gemm == -7 (energy experiment): the sine on the special-function unit
              if constexpr (DBG) sfu_sine = dbg_gemm == -7;
              if (sfu_sine) {
                red2(__ffma2_rn(v, sc2, bb[pr]), r, sx, sy);
                sn = make_float2(__uint_as_float(__float_as_uint(__sinf(r.x)) ^ sx),
                                 __uint_as_float(__float_as_uint(__sinf(r.y)) ^ sy));
              } else {
                sin_red2(__ffma2_rn(v, sc2, bb[pr]), sn, r, sx, sy);
              }
              o[pr] = __fmul2_rn(sn, bc2(A_SCALE));
              tw[2 * pr] = tape_word(r.x, sx);
              tw[2 * pr + 1] = tape_word(r.y, sy);
            }
            publish(kb, o, kst);
            if (wst && (kb == 0 || kb == NKB - 1)) wst[kb == 0 ? 1 : 2] = clock64();
            // global traffic right after the hand-off fence (which waits for everything in flight)
            bool tape_on = !fwd_only;
            if constexpr (DBG) tape_on = tape_on && dbg_gemm != -5;   // -5: energy experiment without tape traffic
            if (tape_on) {
              float4* dst = st + (size_t)(kb * 8) * TM;
              dst[0] = make_float4(tw[0], tw[1], tw[2], tw[3]);
              dst[TM] = make_float4(tw[4], tw[5], tw[6], tw[7]);
            }
            if (kb + 1 < NKB) {
              nb0 = __ldg(bw4 + (kb + 1) * 8);
              nb1 = __ldg(bw4 + (kb + 1) * 8 + 1);
            } else {   // k-block 0 of the next layer (l + 1 <= L exists: this branch is l < L)
              nb0 = ldnc_now(bw4 + H / 4);
              nb1 = ldnc_now(bw4 + H / 4 + 1);
            }
            ISO_STAMP(kst, 5);
          };
          float4 ba0 = pf0, ba1 = pf1, bb0, bb1;
#pragma unroll 1
          for (int kb = 0; kb < NKB; kb += 2) {
            fwd_block(kb, rn, rm, ba0, ba1, bb0, bb1);
            fwd_block(kb + 1, rm, rn, bb0, bb1, ba0, ba1);
          }
          pf0 = ba0;
          pf1 = ba1;
          pf2 = pf3 = make_float4(0.f, 0.f, 0.f, 0.f);   // dead here: do not carry them through the loop above
        } else if (fwd) {
          // ---- E_f(L): sdf = h_L . w_last + b_last ; A = gl_scale * w_last * c_L ----
          const float2 sc2 = bc2(wsi * A_SCALE_INV * omega);
          const float gls = gl_scale * omega;
          const float4* bw4 = reinterpret_cast<const float4*>(biasw + (l - 1) * H) + cslice * 2;
          float2 acc2 = bc2(0.f);
          float mloc = 0.f;
          auto last_block = [&](const int kb, uint32_t(&cur)[8], uint32_t(&nxt)[8]) {
            tmem_ld_wait(cur);
            if (kb + 1 < NKB) tmem_ld8_issue(tacc + (kb + 1) * KB, nxt);
            const float4 b0 = kb ? __ldg(bw4 + kb * 8) : pf0, b1 = kb ? __ldg(bw4 + kb * 8 + 1) : pf1;
            const float4 w0 = __ldg(w_last4 + kb * 8), w1 = __ldg(w_last4 + kb * 8 + 1);
            const float2 bb[4] = {make_float2(b0.x, b0.y), make_float2(b0.z, b0.w), make_float2(b1.x, b1.y),
                                  make_float2(b1.z, b1.w)};
            const float2 ww[4] = {make_float2(w0.x, w0.y), make_float2(w0.z, w0.w), make_float2(w1.x, w1.y),
                                  make_float2(w1.z, w1.w)};
            float2 o[4];
#pragma unroll
            for (int pr = 0; pr < 4; ++pr) {
              const float2 v = make_float2(__uint_as_float(cur[2 * pr]), __uint_as_float(cur[2 * pr + 1]));
              float2 sn, r;
              uint32_t sx, sy;
              sin_red2(__ffma2_rn(v, sc2, bb[pr]), sn, r, sx, sy);
              acc2 = __ffma2_rn(sn, ww[pr], acc2);
              if (!fwd_only) {
                o[pr] = __fmul2_rn(__fmul2_rn(cos_of_reduced(r), signed_scale(gls, sx, sy)), ww[pr]);
                mloc = fmaxf(mloc, fmaxf(fabsf(o[pr].x), fabsf(o[pr].y)));
              }
            }
            if (!fwd_only) publish(kb, o);   // forward-only: no reverse GEMM follows, the next A is the next tile's
          };
#pragma unroll 1
          for (int kb = 0; kb < NKB; kb += 2) {
            last_block(kb, rn, rm);
            last_block(kb + 1, rm, rn);
          }
          if (!fwd_only && L > 1) {   // tape words of k-blocks 0 / 1 of the first reverse stage (layer L - 1)
            const float4* stn = stash + (size_t)(L - 2) * 64 * TM + row + (size_t)(cslice * 2) * TM;
            pf0 = ldcg_now(stn); pf1 = ldcg_now(stn + TM);
            pf2 = ldcg_now(stn + (size_t)8 * TM); pf3 = ldcg_now(stn + (size_t)9 * TM);
          }
          const float acc_sdf = acc2.x + acc2.y;
          row_scale_inv = gl_scale_inv;
          if (cslice) xch[row * 4 + cslice] = acc_sdf;
          row_barrier(q);
          if (cslice == 0) {
            sdf_row = ((acc_sdf + xch[row * 4 + 1]) + (xch[row * 4 + 2] + xch[row * 4 + 3])) + b_last;
            if (grow < n && sdf_out) sdf_out[grow] = sdf_row;
          }
          row_barrier(q);
          if (!fwd_only) {   // row maximum of what was just written: the next stage derives its scale from it
            xch[row * 4 + cslice] = mloc;
            row_barrier(q);
            row_max = fmaxf(fmaxf(xch[row * 4], xch[row * 4 + 1]), fmaxf(xch[row * 4 + 2], xch[row * 4 + 3]));
            row_barrier(q);
          }
          if (nw.mode == 1 && cslice == 0) {
            // ---- fused ray-marching step on this row (same arithmetic as trace_step_kernel, project.cu) ----
            bool still = false;
            int p = 0;
            float nx = 0.f, ny = 0.f, nz = 0.f;
            if (grow < n) {
              p = nw.act_in ? nw.act_in[grow] : grow;
              nw.eval[p] = sdf_row;
              still = fabsf(sdf_row) > nw.tol;
              if (still) {
                nx = px; ny = py; nz = pz;   // the evaluated position IS points[p]
                if (nw.do_update) {
                  const float af = __fmul_rn(nw.alpha, sdf_row);
                  const float mx = __fmul_rn(af, nw.dirs[3 * (size_t)p]);
                  const float my = __fmul_rn(af, nw.dirs[3 * (size_t)p + 1]);
                  const float mz = __fmul_rn(af, nw.dirs[3 * (size_t)p + 2]);
                  const float nrm = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(mx, mx), __fmul_rn(my, my)), __fmul_rn(mz, mz)));
                  const float dn = fmaxf(nrm, 1e-15f);
                  const float len = fminf(nrm, nw.max_step);
                  nx = __fadd_rn(nx, __fmul_rn(__fdiv_rn(mx, dn), len));
                  ny = __fadd_rn(ny, __fmul_rn(__fdiv_rn(my, dn), len));
                  nz = __fadd_rn(nz, __fmul_rn(__fdiv_rn(mz, dn), len));
                  const float r = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(nx, nx), __fmul_rn(ny, ny)), __fmul_rn(nz, nz)));
                  if (r < nw.bound) {
                    nw.points[3 * (size_t)p] = nx;
                    nw.points[3 * (size_t)p + 1] = ny;
                    nw.points[3 * (size_t)p + 2] = nz;
                  } else {
                    still = false;               // left the sphere: old position stays, the ray retires
                  }
                }
              }
            }
            append_active(still, p, nx, ny, nz);
          }
        } else if (l > 1) {
          // ---- E_b(l): g_{l-1} = acc / scales ; gp_{l-1} = g_{l-1} * w cos(w z_{l-1}) -> A (row-scaled) ----
          const float sc = wsi * row_scale_inv;
          // Per-row scale of the next A operand (gp_{l-1}) WITHOUT a pass over the accumulator: the rows written
          // one stage ago had the measured maximum `row_max` (scaled), so max_j |g_{l-1,j}| |omega| <=
          // (row_max / row scale) * |omega| max_j sum_k |W_l[k][j]| (header gain).  The bound is a few binades
          // above the true maximum (the sum does not see the cancellation); the fp16 hi/lo pair keeps 22 bits of
          // the row maximum anywhere in [2^-3, 2^15], so landing at 2^7..2^10 instead of 2^11 costs nothing, and
          // the next stage re-centres on the maximum it measures itself -- the slack does not accumulate.
          const float new_scale = pow2_scale_for(row_max * row_scale_inv * shdr[HDR_GAIN + l - 1]);
          float mloc = 0.f;
          const float2 scs2 = bc2(sc * new_scale * omega);   // the tape holds cos without its factor omega
          const float4* st = stash + (size_t)(l - 2) * 64 * TM + row + (size_t)(cslice * 2) * TM;
          // one k-block: (t0, t1) = its 8 tape words, replaced at the end by those of k-block kb + 2 (an L2 hit is
          // ~1 k-block of this loop away, a miss more); those of k-blocks 0 and 1 were requested one stage ago
          auto rev_block = [&](const int kb, uint32_t(&cur)[8], uint32_t(&nxt)[8], float4& t0, float4& t1) {
            long long* kst = nullptr;
            if constexpr (DBG) kst = kbase ? kbase + kb * 8 : nullptr;
            ISO_STAMP(kst, 0);
            tmem_ld_wait(cur);
            ISO_STAMP(kst, 1);
            if (kb + 1 < NKB) tmem_ld8_issue(tacc + (kb + 1) * KB, nxt);
            float2 o[4];
            o[0] = __fmul2_rn(__fmul2_rn(make_float2(__uint_as_float(cur[0]), __uint_as_float(cur[1])), scs2),
                              cos_of_tape(t0.x, t0.y));
            o[1] = __fmul2_rn(__fmul2_rn(make_float2(__uint_as_float(cur[2]), __uint_as_float(cur[3])), scs2),
                              cos_of_tape(t0.z, t0.w));
            o[2] = __fmul2_rn(__fmul2_rn(make_float2(__uint_as_float(cur[4]), __uint_as_float(cur[5])), scs2),
                              cos_of_tape(t1.x, t1.y));
            o[3] = __fmul2_rn(__fmul2_rn(make_float2(__uint_as_float(cur[6]), __uint_as_float(cur[7])), scs2),
                              cos_of_tape(t1.z, t1.w));
            mloc = fmaxf(fmaxf(mloc, fmaxf(fabsf(o[0].x), fabsf(o[0].y))), fmaxf(fabsf(o[1].x), fabsf(o[1].y)));
            mloc = fmaxf(fmaxf(mloc, fmaxf(fabsf(o[2].x), fabsf(o[2].y))), fmaxf(fabsf(o[3].x), fabsf(o[3].y)));
            publish(kb, o, kst);
            if (wst && (kb == 0 || kb == NKB - 1)) wst[kb == 0 ? 1 : 2] = clock64();
            // this k-block's tape lines are dead now (the next tile rewrites them in full before reading):
            // drop them from L2 instead of letting them be written back to HBM
            if ((lane & 7) == 0) {
              asm volatile("discard.global.L2 [%0], 128;" ::"l"(st + (size_t)(kb * 8) * TM) : "memory");
              asm volatile("discard.global.L2 [%0], 128;" ::"l"(st + (size_t)(kb * 8 + 1) * TM) : "memory");
            }
            bool tape_on = true;
            if constexpr (DBG) tape_on = dbg_gemm != -5;
            if (kb + 2 < NKB && tape_on) {   // right after the hand-off fence, two k-blocks ahead of its use
              t0 = __ldcg(st + (size_t)((kb + 2) * 8) * TM);
              t1 = __ldcg(st + (size_t)((kb + 2) * 8 + 1) * TM);
            }
            ISO_STAMP(kst, 5);
          };
          float4 e0 = pf0, e1 = pf1, f0 = pf2, f1 = pf3;
#pragma unroll 1
          for (int kb = 0; kb < NKB; kb += 2) {
            rev_block(kb, rn, rm, e0, e1);
            rev_block(kb + 1, rm, rn, f0, f1);
          }
          if (l > 2) {   // k-blocks 0 / 1 of the next reverse stage (layer l - 2's tape)
            const float4* stn = stash + (size_t)(l - 3) * 64 * TM + row + (size_t)(cslice * 2) * TM;
            pf0 = ldcg_now(stn); pf1 = ldcg_now(stn + TM);
            pf2 = ldcg_now(stn + (size_t)8 * TM); pf3 = ldcg_now(stn + (size_t)9 * TM);
          }
          row_scale_inv = 1.f / new_scale;
          xch[row * 4 + cslice] = mloc;      // off the critical path: the next accumulator is ~2 k cycles away
          row_barrier(q);
          row_max = fmaxf(fmaxf(xch[row * 4], xch[row * 4 + 1]), fmaxf(xch[row * 4 + 2], xch[row * 4 + 3]));
          row_barrier(q);
        } else {
          // ---- E_b(1): g_0 = acc / scales ; gp_0 = g_0 * w0 cos(w0 z_0) ; grad = gp_0 W_0 ----
          // (the w0 table holds omega_0-scaled rows, so gp_0 . W_0 = sum (g_0 cos) * (omega_0 W_0))
          const float sc = wsi * row_scale_inv;
          float2 gx2 = bc2(0.f), gy2 = bc2(0.f), gz2 = bc2(0.f);
          auto first_block = [&](const int kb, uint32_t(&cur)[8], uint32_t(&nxt)[8]) {
            tmem_ld_wait(cur);
            if (kb + 1 < NKB) tmem_ld8_issue(tacc + (kb + 1) * KB, nxt);
#pragma unroll
            for (int pr = 0; pr < 4; ++pr) {
              const float2 v = make_float2(__uint_as_float(cur[2 * pr]), __uint_as_float(cur[2 * pr + 1]));
              const float4 wa = __ldg(w0p + kb * 32 + pr * 2), wb = __ldg(w0p + kb * 32 + pr * 2 + 1);
              const float2 wx = make_float2(wa.x, wa.y), wy = make_float2(wa.z, wa.w), wz = make_float2(wb.x, wb.y);
              const float2 th = __ffma2_rn(wz, pz2, __ffma2_rn(wy, py2, __ffma2_rn(wx, px2, make_float2(wb.z, wb.w))));
              float2 r;
              uint32_t sx, sy;
              red2(th, r, sx, sy);   // the layer-0 cosine is recomputed, not taped: its argument costs 3 FFMA2
              const float2 gp = __fmul2_rn(__fmul2_rn(v, signed_scale(sc, sx, sy)), cos_of_reduced(r));
              gx2 = __ffma2_rn(gp, wx, gx2);
              gy2 = __ffma2_rn(gp, wy, gy2);
              gz2 = __ffma2_rn(gp, wz, gz2);
            }
          };
#pragma unroll 1
          for (int kb = 0; kb < NKB; kb += 2) {
            first_block(kb, rn, rm);
            first_block(kb + 1, rm, rn);
          }
          const float gx = gx2.x + gx2.y, gy = gy2.x + gy2.y, gz = gz2.x + gz2.y;
          tc_fence_before();
          // The slot was last written as an A chunk by another warp of this row group two stages ago; that
          // write is ordered before this one through the a_ready / acc_full mbarrier chain (the GEMM that
          // consumed it has completed).  The extra bar.sync only restates that ordering in a form
          // compute-sanitizer's racecheck tracks (~100 cycles per tile).
          row_barrier(q);
          if (cslice) *reinterpret_cast<float4*>(gsc + (cslice - 1) * A_LBO) = make_float4(gx, gy, gz, 0.f);
          row_barrier(q);
          float fgx = 0.f, fgy = 0.f, fgz = 0.f;
          if (cslice == 0) {
            const float4 d0 = *reinterpret_cast<const float4*>(gsc);
            const float4 d1 = *reinterpret_cast<const float4*>(gsc + A_LBO);
            const float4 d2 = *reinterpret_cast<const float4*>(gsc + 2 * A_LBO);
            fgx = (gx + d0.x) + (d1.x + d2.x);
            fgy = (gy + d0.y) + (d1.y + d2.y);
            fgz = (gz + d0.z) + (d1.z + d2.z);
            if (grow < n && grad_out) {
              grad_out[3 * (size_t)grow] = fgx;
              grad_out[3 * (size_t)grow + 1] = fgy;
              grad_out[3 * (size_t)grow + 2] = fgz;
            }
          }
          row_barrier(q);
          if (nw.points && cslice == 0) {
            // ---- fused Newton step on this row (warps 0..3 hold the tile's 128 rows) ----
            bool still = false;
            int p = 0;
            float nx = 0.f, ny = 0.f, nz = 0.f;
            if (grow < n) {
              p = nw.act_in ? nw.act_in[grow] : grow;
              nw.normals[3 * (size_t)p] = fgx;
              nw.normals[3 * (size_t)p + 1] = fgy;
              nw.normals[3 * (size_t)p + 2] = fgz;
              still = fabsf(sdf_row) > nw.tol;
              nw.not_conv[p] = still ? 1 : 0;
              if (still) {
                nx = px; ny = py; nz = pz;   // the evaluated position IS points[p]
                if (nw.do_update) {
                  const float ss = __fadd_rn(__fadd_rn(__fmul_rn(fgx, fgx), __fmul_rn(fgy, fgy)), __fmul_rn(fgz, fgz));
                  const float den = eps_denom_f(ss, 1.0e-17f);
                  const float mx = __fmul_rn(sdf_row, __fdiv_rn(fgx, den));
                  const float my = __fmul_rn(sdf_row, __fdiv_rn(fgy, den));
                  const float mz = __fmul_rn(sdf_row, __fdiv_rn(fgz, den));
                  const float nrm = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(mx, mx), __fmul_rn(my, my)), __fmul_rn(mz, mz)));
                  const float dn = fmaxf(nrm, 1e-15f);           // F.normalize(eps=1e-15)
                  const float len = fminf(nrm, nw.max_step);     // clamp_max(0.1)
                  nx = __fsub_rn(nx, __fmul_rn(__fdiv_rn(mx, dn), len));
                  ny = __fsub_rn(ny, __fmul_rn(__fdiv_rn(my, dn), len));
                  nz = __fsub_rn(nz, __fmul_rn(__fdiv_rn(mz, dn), len));
                  nw.points[3 * (size_t)p] = nx;
                  nw.points[3 * (size_t)p + 1] = ny;
                  nw.points[3 * (size_t)p + 2] = nz;
                }
              }
            }
            append_active(still, p, nx, ny, nz);
          }
        }
        ISO_STAMP(tstamp, 2);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == N_EPI_WARPS + 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
  // every CTA has taken its last (out-of-range) ticket before it gets here: the last one re-arms the slot
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(sched + 1, 1) == (int)gridDim.x - 1) {
      sched[0] = 0;
      sched[1] = 0;
      __threadfence();
    }
  }
}

}  // namespace siren
}  // namespace isob200

using namespace isob200;
using namespace isob200::siren;

extern "C" {

size_t isob200_siren_blob_bytes(int n_hidden) {
  if (n_hidden < 1 || n_hidden > MAX_LAYERS) return 0;
  return blob_bytes(n_hidden);
}
size_t isob200_siren_pack_ws_bytes(void) { return PACK_WS_BYTES; }
size_t isob200_siren_scratch_bytes(int n_hidden) {
  int l = n_hidden > 1 ? n_hidden - 1 : 1;
  return (size_t)kNumSMs * l * 64 * TM * sizeof(float4);
}

int isob200_siren_pack(const float* w0, const float* b0, const float* w_hidden, const float* b_hidden,
                       const float* w_last, const float* b_last, float omega0, float omega, int hidden,
                       int n_hidden, void* blob, size_t blob_bytes_, void* ws, size_t ws_bytes, void* stream) {
  ISO_CHECK_ARG(hidden == H, "siren_pack: only hidden width %d is built (got %d)", H, hidden);
  ISO_CHECK_ARG(n_hidden >= 1 && n_hidden <= MAX_LAYERS, "siren_pack: n_hidden must be in [1, %d]", MAX_LAYERS);
  ISO_CHECK_ARG(w0 && w_hidden && w_last && blob && ws, "siren_pack: null pointer");
  ISO_CHECK_ARG(blob_bytes_ >= blob_bytes(n_hidden), "siren_pack: blob too small");
  if (ws_bytes < PACK_WS_BYTES) {
    set_error("siren_pack: workspace too small");
    return ISOB200_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  ISO_CUDA(cudaMemsetAsync(ws, 0, PACK_WS_BYTES, st));
  siren_absmax_kernel<<<dim3(32, n_hidden + 1), 256, 0, st>>>(w_hidden, w_last, n_hidden, (unsigned*)ws);
  ISO_CHECK_LAUNCH("siren_absmax_kernel");
  siren_pack_kernel<<<kNumSMs * 2, 256, 0, st>>>(w0, b0, w_hidden, b_hidden, w_last, b_last, omega0, omega, n_hidden,
                                                (const unsigned*)ws, (unsigned char*)blob);
  ISO_CHECK_LAUNCH("siren_pack_kernel");
  siren_gain_kernel<<<n_hidden, H, 0, st>>>(w_hidden, omega, (unsigned char*)blob);
  ISO_CHECK_LAUNCH("siren_gain_kernel");
  return ISOB200_OK;
}

static int g_siren_max_ctas = kNumSMs;   // tuning knob: persistent CTAs per launch (<= one per SM)
static int g_siren_stagger = 0;          // tuning knob: start delay of the CTAs on odd SMs in cycles (-1 = half a tile)
static int g_siren_stagger_min_tiles = 3 * kNumSMs;   // ... for launches with at least this many tiles
static std::atomic<unsigned> g_siren_launch_seq{0};   // ctypes releases the GIL: host threads may launch concurrently
int isob200_siren_set_stagger(int cycles, int min_tiles) {
  const int old = g_siren_stagger;
  g_siren_stagger = cycles;
  if (min_tiles > 0) g_siren_stagger_min_tiles = min_tiles;
  return old;
}
int isob200_siren_set_max_ctas(int n) {
  const int old = g_siren_max_ctas;
  g_siren_max_ctas = n < 1 ? 1 : (n > kNumSMs ? kNumSMs : n);
  return old;
}

static int launch_siren(const float* x, int n_max, const int* n_dev, const void* blob, int n_hidden, float* sdf,
                        float* grad, void* scratch, size_t scratch_bytes, float* dbg, int dbg_gemm,
                        const Newton& nw, void* stream, const char* who) {
  ISO_CHECK_ARG(n_hidden >= 1 && n_hidden <= MAX_LAYERS, "%s: n_hidden must be in [1, %d]", who, MAX_LAYERS);
  ISO_CHECK_ARG(n_max >= 0, "%s: negative point count", who);
  if (n_max == 0) return ISOB200_OK;
  ISO_CHECK_ARG(x && blob && scratch, "%s: null pointer", who);
  if (scratch_bytes < isob200_siren_scratch_bytes(n_hidden)) {
    set_error("%s: scratch too small", who);
    return ISOB200_ERR_WORKSPACE;
  }
  {
    // opt-in to > 48 KB of dynamic shared memory: once per device
    static bool attr_done[64] = {};
    int dev = 0;
    ISO_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || !attr_done[dev]) {
      ISO_CUDA(cudaFuncSetAttribute(siren_sdf_grad_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
      ISO_CUDA(cudaFuncSetAttribute(siren_sdf_grad_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
      if (dev >= 0 && dev < 64) attr_done[dev] = true;
    }
  }
  int tiles = (n_max + TM - 1) / TM;
  int grid = tiles < g_siren_max_ctas ? tiles : g_siren_max_ctas;
  Newton nwk = nw;
  nwk.sched_slot = (int)(g_siren_launch_seq.fetch_add(1, std::memory_order_relaxed) % SCHED_SLOTS);
  // half a tile period: a tile is 2 n_hidden stages of ~10.5 k cycles plus ~20 k for the two SIMT-only ends
  nwk.stagger_cycles = g_siren_stagger >= 0 ? g_siren_stagger : (2 * n_hidden * 10500 + 20000) / 2;
  nwk.stagger_min_tiles = g_siren_stagger_min_tiles;
  if (dbg)
    siren_sdf_grad_kernel<true><<<grid, THREADS, SMEM_BYTES, (cudaStream_t)stream>>>(
        x, n_max, n_dev, (const unsigned char*)blob, n_hidden, sdf, grad, (float*)scratch, dbg, dbg_gemm, nwk);
  else
    siren_sdf_grad_kernel<false><<<grid, THREADS, SMEM_BYTES, (cudaStream_t)stream>>>(
        x, n_max, n_dev, (const unsigned char*)blob, n_hidden, sdf, grad, (float*)scratch, nullptr, -1, nwk);
  ISO_CHECK_LAUNCH("siren_sdf_grad_kernel");
  return ISOB200_OK;
}

int isob200_siren_sdf_grad(const float* x, int n_max, const int* n_dev, const void* blob, int n_hidden, float* sdf,
                           float* grad, void* scratch, size_t scratch_bytes, float* dbg, int dbg_gemm,
                           void* stream) {
  ISO_CHECK_ARG(n_max == 0 || (sdf && grad), "siren_sdf_grad: null output");
  Newton nw = {};
  return launch_siren(x, n_max, n_dev, blob, n_hidden, sdf, grad, scratch, scratch_bytes, dbg, dbg_gemm, nw, stream,
                      "siren_sdf_grad");
}

// One Newton iteration of _project_points with the SDF evaluation fused in: isob200_siren_sdf_grad at
// x (the positions of the active rows) followed by isob200_project_step, in one kernel.  Differences
// from that pair: act_out / next_points are appended per tile in completion order (rows are independent,
// so the results are identical), and *count_out must be zero before the call.
int isob200_siren_project_step(const float* x, int n_max, const int* n_dev, const void* blob, int n_hidden,
                               void* scratch, size_t scratch_bytes, float* points, float* normals,
                               unsigned char* not_converged, const int* act_in, float tol, float max_step,
                               int do_update, int* act_out, float* next_points, int* count_out, void* stream) {
  ISO_CHECK_ARG(n_max == 0 || (points && normals && not_converged && act_out && count_out),
                "siren_project_step: null pointer");
  ISO_CHECK_ARG(n_dev != count_out, "siren_project_step: n_dev must not alias count_out");
  Newton nw = {points, normals, not_converged, act_in, act_out, next_points, count_out, tol, max_step, do_update};
  return launch_siren(x, n_max, n_dev, blob, n_hidden, nullptr, nullptr, scratch, scratch_bytes, nullptr, -1, nw,
                      stream, "siren_project_step");
}

// SDF value only: the forward half of isob200_siren_sdf_grad (half the GEMMs, no tape).  Bit-identical to the
// sdf output of isob200_siren_sdf_grad.
int isob200_siren_sdf(const float* x, int n_max, const int* n_dev, const void* blob, int n_hidden, float* sdf,
                      void* scratch, size_t scratch_bytes, void* stream) {
  ISO_CHECK_ARG(n_max == 0 || sdf, "siren_sdf: null output");
  Newton nw = {};
  nw.mode = 2;
  return launch_siren(x, n_max, n_dev, blob, n_hidden, sdf, nullptr, scratch, scratch_bytes, nullptr, -1, nw, stream,
                      "siren_sdf");
}

// One iteration of SphereTracing.project_points (levelset_sampling.py:733-786) with the SDF evaluation
// fused in: the FORWARD half of the SIREN kernel only (the march needs no gradient, so half the GEMMs and
// no tape) followed by the update rule of isob200_trace_step.  Conventions as isob200_siren_project_step.
int isob200_siren_trace_step(const float* x, int n_max, const int* n_dev, const void* blob, int n_hidden,
                             void* scratch, size_t scratch_bytes, float* points, const float* dirs, float* eval,
                             const int* act_in, float active_tol, float alpha, float max_step, float bound,
                             int do_update, int* act_out, float* next_points, int* count_out, void* stream) {
  ISO_CHECK_ARG(n_max == 0 || (points && dirs && eval && act_out && count_out), "siren_trace_step: null pointer");
  ISO_CHECK_ARG(n_dev != count_out, "siren_trace_step: n_dev must not alias count_out");
  Newton nw = {points, nullptr, nullptr, act_in, act_out, next_points, count_out, active_tol, max_step, do_update,
               1, dirs, eval, alpha, bound};
  return launch_siren(x, n_max, n_dev, blob, n_hidden, nullptr, nullptr, scratch, scratch_bytes, nullptr, -1, nw,
                      stream, "siren_trace_step");
}

}  // extern "C"
