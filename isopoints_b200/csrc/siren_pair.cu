// Fused SIREN SDF + input-gradient kernel on CTA PAIRS (tcgen05 cta_group::2), two tiles in flight.
//
// Same math, packing, precision scheme and results as siren.cu (see its header); what changes is the
// mapping, to break that kernel's latency chain (a tile's 2L GEMMs and their epilogues are strictly
// sequential, so with one tile per SM the tensor pipe and the SIMT pipes each idle half of the time):
//   * a tile of 128 points is shared by the two CTAs of a cluster: tcgen05.mma.cta_group::2 with M = 128
//     spans the pair, each CTA holding 64 rows of A (64 KB as fp16 hi | lo instead of 128 KB) and the
//     128 weight rows of its rank (16 KB stages, half the shared-memory reads and half the TMA traffic per
//     SM); each SM's accumulator is 128 TMEM columns (rows on lanes 0-63 hold output columns 0-127, lanes
//     64-127 hold columns 128-255; layout pinned by csrc/umma2_probe.cu);
//   * that halving leaves room for TWO tiles in flight per pair (2 x 64 KB of A, 4 x 128 accumulator
//     columns): the 16 epilogue warps alternate between the two tiles stage by stage, so one tile's
//     tensor-core work runs under the other's sin / cos / fp16-split epilogue, and a lone tile (late Newton
//     iterations) walks its chain with half the epilogue work per SM.
// Roles per CTA: warps 0-15 epilogue (TMEM lane quarter q = warp % 4: row = 32 (q & 1) + lane, column half
// q >> 1; 8-column slice warp / 4 of each k-block), warp 16 weight producer (1-D TMA bulk copies of this
// rank's stages), warp 17: TMEM allocation; lane 0 issues every MMA on the leader CTA and relays
// "my stage has landed" to the leader on the other one.  Hand-offs that cross the pair use remote
// mbarrier arrives (mapa) and multicast tcgen05.commit.
#include "siren_common.cuh"

namespace isob200 {
namespace siren {
namespace pr {

constexpr int HR = 64;                         // rows of a tile held by one CTA
constexpr int PA_LBO = HR * 16;                // 1024: K-chunk stride of the A half-tile
constexpr int PA_PART = HR * H * 2;            // 32 KB
constexpr int PA_SLOT = 2 * PA_PART;           // 64 KB: hi | lo
constexpr int PB_LBO = 128 * 16;               // 2048: K-chunk stride of a B half-stage
constexpr int PB_PART = 128 * KB * 2;          // 8 KB
constexpr int PSTAGE = 2 * PB_PART;            // 16 KB
constexpr int NST = 6;                         // weight stages in flight
constexpr int PSM_STAGE = 2 * PA_SLOT;                    // 131072
constexpr int PSM_XCH = PSM_STAGE + NST * PSTAGE;         // 229376: float xch[64][8]
constexpr int PSM_BAR = PSM_XCH + HR * 8 * 4;             // 215040
constexpr int PB_AREADY = 0;                   // [2 slots][8 k-blocks]  (used on the leader)
constexpr int PB_ACCFULL = 16;                 // [2 slots][2 buffers]
constexpr int PB_WFULL = 20;                   // [NST]
constexpr int PB_WPEER = 26;                   // [NST]                  (used on the leader)
constexpr int PB_WEMPTY = 32;                  // [NST]
constexpr int PSM_TMEM = PSM_BAR + 40 * 8;
constexpr int PSM_CMP = PSM_TMEM + 16;
constexpr int PSMEM_BYTES = PSM_CMP + 32;      // 215408

// bring-up / tuning aid: cycle stamps of pair 0's MMA issuer and of its leader's epilogue warp 0 (first 64 GEMMs)
__device__ long long g_pair_stamps[64 * 8];

// k-block visiting order: the two column halves of the epilogue produce k-blocks j and 4 + j together
__device__ __forceinline__ int kbo(int i) { return (i >> 1) + ((i & 1) << 2); }

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1)
siren_pair_kernel(const float* __restrict__ x, int n_max, const int* __restrict__ n_dev,
                  const unsigned char* __restrict__ blob, int L, float* __restrict__ sdf_out,
                  float* __restrict__ grad_out, float* __restrict__ scratch, const Newton nw) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t sbase = smem_u32(smem);
  const uint32_t rank = cluster_ctarank();
  auto bar = [&](int i) { return sbase + PSM_BAR + 8 * i; };

  int n = n_max;
  if (n_dev) {
    int nd = *n_dev;
    n = nd < n_max ? nd : n_max;
  }
  const int T = (n + TM - 1) / TM;          // tiles of 128 rows
  const int P = blockIdx.x >> 1;            // this pair
  const int NP = gridDim.x >> 1;
  const int nj = T > P ? (T - P + NP - 1) / NP : 0;   // tiles of this pair: P, P + NP, ...
  const int rounds = (nj + 1) >> 1;                   // two tiles (slots 0, 1) per round

  if (threadIdx.x == 0) {
    for (int i = 0; i < 16; ++i) mbar_init(bar(PB_AREADY + i), 16);   // 8 warps of each CTA
    for (int i = 0; i < 4; ++i) mbar_init(bar(PB_ACCFULL + i), 1);
    for (int i = 0; i < NST; ++i) {
      mbar_init(bar(PB_WFULL + i), 1);
      mbar_init(bar(PB_WPEER + i), 1);
      mbar_init(bar(PB_WEMPTY + i), 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  cluster_sync_all();
  if (warp == N_EPI_WARPS + 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sbase + PSM_TMEM), "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem + PSM_TMEM);

  const float* hdr = reinterpret_cast<const float*>(blob);
  const size_t img0 = off_images_split(L);
  const int n_gemm = 2 * L;

  if (warp == N_EPI_WARPS) {
    // ===================== weight producer (this rank's half of every stage) =====================
    if (lane == 0) {
      uint32_t it = 0;
      for (int rd = 0; rd < rounds; ++rd) {
        const int ns = (2 * rd + 1 < nj) ? 2 : 1;
        for (int g = 0; g < n_gemm; ++g) {
          const int l = g < L ? g : (2 * L - 1 - g);
          const int o = g < L ? 0 : 1;
          const unsigned char* img = blob + img0 + (size_t)(l * 2 + o) * image_bytes();
          for (int s = 0; s < ns; ++s) {
            for (int i = 0; i < NKB; ++i, ++it) {
              const int kb = kbo(i);
              const uint32_t st = it % NST, ph = (it / NST) & 1;
              mbar_wait(bar(PB_WEMPTY + st), ph ^ 1);
              mbar_expect_tx(bar(PB_WFULL + st), PSTAGE);
              tma_bulk_g2s(sbase + PSM_STAGE + st * PSTAGE, img + (size_t)(kb * 2 + rank) * PSTAGE, PSTAGE,
                           bar(PB_WFULL + st));
            }
          }
        }
      }
    }
  } else if (warp == N_EPI_WARPS + 1) {
    if (lane == 0 && rank != 0) {
      // ===================== relay: tell the leader that this CTA's half of a stage has landed ==========
      uint32_t it = 0;
      for (int rd = 0; rd < rounds; ++rd) {
        const int ns = (2 * rd + 1 < nj) ? 2 : 1;
        for (int c = 0; c < n_gemm * ns * NKB; ++c, ++it) {
          const uint32_t st = it % NST, ph = (it / NST) & 1;
          mbar_wait(bar(PB_WFULL + st), ph);
          mbar_arrive_cluster_relaxed(bar(PB_WPEER + st), 0);
        }
      }
    } else if (lane == 0) {
      // ===================== MMA issuer (leader CTA; every MMA spans the pair) =====================
      uint32_t it = 0;
      uint32_t Gs[2] = {0, 0};
      for (int rd = 0; rd < rounds; ++rd) {
        const int ns = (2 * rd + 1 < nj) ? 2 : 1;
        for (int g = 0; g < n_gemm; ++g) {
#pragma unroll
          for (int s = 0; s < 2; ++s) {
            if (s >= ns) continue;
            const uint32_t G = Gs[s];
            const uint32_t d_tmem = tmem_base + (uint32_t)(s * 2 + (G & 1)) * 128;
            const uint32_t gi = Gs[0] + Gs[1];
            const bool stamp = blockIdx.x == 0 && gi < 64;
            long long wa = 0, ww = 0, wp = 0, t_first = 0;
            for (int i = 0; i < NKB; ++i, ++it) {
              const int kb = kbo(i);
              const uint32_t st = it % NST, ph = (it / NST) & 1;
              const long long t0 = stamp ? clock64() : 0;
              mbar_wait(bar(PB_AREADY + s * 8 + kb), G & 1);
              const long long t1 = stamp ? clock64() : 0;
              mbar_wait(bar(PB_WFULL + st), ph);
              const long long t2 = stamp ? clock64() : 0;
              mbar_wait(bar(PB_WPEER + st), ph);
              if (stamp) {
                const long long t3 = clock64();
                wa += t1 - t0; ww += t2 - t1; wp += t3 - t2;
                if (i == 0) t_first = t1;
              }
              tc_fence_after();
              const uint32_t a_hi = sbase + s * PA_SLOT + kb * (KB / 8) * PA_LBO;
              const uint32_t a_lo = a_hi + PA_PART;
              const uint32_t b_hi = sbase + PSM_STAGE + st * PSTAGE;
              const uint32_t b_lo = b_hi + PB_PART;
#pragma unroll
              for (int k16 = 0; k16 < KB / 16; ++k16) {
                const uint64_t dah = make_desc(a_hi + k16 * 2 * PA_LBO, PA_LBO);
                const uint64_t dal = make_desc(a_lo + k16 * 2 * PA_LBO, PA_LBO);
                const uint64_t dbh = make_desc(b_hi + k16 * 2 * PB_LBO, PB_LBO);
                const uint64_t dbl = make_desc(b_lo + k16 * 2 * PB_LBO, PB_LBO);
                tc_mma2_f16(d_tmem, dal, dbh, IDESC, (i | k16) ? 1u : 0u);
                tc_mma2_f16(d_tmem, dah, dbl, IDESC, 1u);
                tc_mma2_f16(d_tmem, dah, dbh, IDESC, 1u);
              }
              tc_commit2(bar(PB_WEMPTY + st));   // both producers may refill this stage
            }
            tc_commit2(bar(PB_ACCFULL + s * 2 + (G & 1)));   // both CTAs: accumulator complete
            if (stamp) {
              long long* t = g_pair_stamps + gi * 8;
              t[0] = t_first; t[1] = clock64(); t[2] = wa; t[3] = ww; t[4] = wp; t[5] = s;
            }
            Gs[s] = G + 1;
          }
        }
      }
    }
  } else {
    // ===================== epilogue warps =====================
    const int q = warp & 3;            // TMEM lane quarter
    const int w4 = warp >> 2;          // 8-column slice inside every 32-column k-block
    const int ch = q >> 1;             // column half: lanes 64..127 hold output columns 128..255
    const int hrow = 32 * (q & 1) + lane;   // row of the half-tile
    const int pi = ch * 4 + w4;        // which of the row's 8 threads this is; 0 owns the row's results
    const uint32_t tl = tmem_base + ((uint32_t)(32 * q) << 16) + 8 * w4;
    const uint32_t a_thr = (uint32_t)w4 * PA_LBO + (uint32_t)(hrow >> 3) * SBO + (uint32_t)(hrow & 7) * 16;
    float* xch = reinterpret_cast<float*>(smem + PSM_XCH);
    const float4* w0p = reinterpret_cast<const float4*>(blob + OFF_W0B) + w4 * 8;
    const float4* w_last4 = reinterpret_cast<const float4*>(blob + OFF_WLAST) + w4 * 2;
    const float* biasw = reinterpret_cast<const float*>(blob + OFF_BIAS);
    const float omega = hdr[HDR_OMEGA];
    const float gl_scale = hdr[HDR_GL_SCALE], gl_scale_inv = hdr[HDR_GL_SCALE_INV];
    const float b_last = hdr[HDR_B_LAST];
    const int nl1 = L > 1 ? L - 1 : 1;
    // per-CTA stash: [slot][(l-1)][col4 (64)][row (64)] float4
    float4* stash0 = reinterpret_cast<float4*>(scratch) + (size_t)blockIdx.x * 2 * nl1 * 64 * HR;
    // barrier among the 8 warps that share a row group (q & 1): both column halves, four slices
    auto row_barrier8 = [&]() { asm volatile("bar.sync %0, 256;" ::"r"((q & 1) + 1) : "memory"); };

    uint32_t Gs[2] = {0, 0};
    float rsi[2] = {1.f, 1.f};      // inverse scale of this row of the current backward A, per slot
    float sdfr[2] = {0.f, 0.f};     // the row's sdf (owner thread), per slot

    for (int rd = 0; rd < rounds; ++rd) {
      const int ns = (2 * rd + 1 < nj) ? 2 : 1;
      for (int sg = 0; sg <= n_gemm; ++sg) {
#pragma unroll
        for (int s = 0; s < 2; ++s) {
          if (s >= ns) continue;
          const int tile = P + (2 * rd + s) * NP;
          const int grow = tile * TM + (int)rank * HR + hrow;
          const bool valid = grow < n;
          float4* stash = stash0 + (size_t)s * nl1 * 64 * HR;
          // publish this thread's K-chunk of k-block kb of slot s, one (remote) arrive per warp on the leader
          auto publish = [&](int kb, const float2* o) {
            const uint32_t off = (uint32_t)(s * PA_SLOT) + (uint32_t)(kb * 4) * PA_LBO + a_thr;
            store_chunk(sbase + off, sbase + off + PA_PART, o);
            tc_fence_before();
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster_relaxed(bar(PB_AREADY + s * 8 + kb), 0);
          };
          float px = 0.f, py = 0.f, pz = 0.f;
          if ((sg == 0 || sg == n_gemm) && valid) {   // first layer and its recomputation in the last stage
            px = __ldcg(x + 3 * (size_t)grow);
            py = __ldcg(x + 3 * (size_t)grow + 1);
            pz = __ldcg(x + 3 * (size_t)grow + 2);
          }
          const float2 px2 = bc2(px), py2 = bc2(py), pz2 = bc2(pz);

          if (sg == 0) {
            // ---- E0: first layer in SIMT, A = 2^12 sin(w0 z_0) ----
            rsi[s] = 1.f;
#pragma unroll 1
            for (int j = 0; j < 4; ++j) {
              const int kb = 4 * ch + j;
              float2 o[4];
#pragma unroll
              for (int p2 = 0; p2 < 4; ++p2) {
                const float4 wa = __ldg(w0p + kb * 32 + p2 * 2), wb = __ldg(w0p + kb * 32 + p2 * 2 + 1);
                const float2 th = __ffma2_rn(make_float2(wb.x, wb.y), pz2,
                                             __ffma2_rn(make_float2(wa.z, wa.w), py2,
                                                        __ffma2_rn(make_float2(wa.x, wa.y), px2,
                                                                   make_float2(wb.z, wb.w))));
                float2 sn, cp;
                uint32_t sx, sy;
                sincos2(th, sn, cp, sx, sy);
                o[p2] = __fmul2_rn(sn, bc2(A_SCALE));
              }
              publish(kb, o);
            }
            continue;
          }

          const int g = sg - 1;
          const uint32_t G = Gs[s];
          const uint32_t gi = Gs[0] + Gs[1];
          Gs[s] = G + 1;
          const uint32_t buf = G & 1;
          mbar_wait(bar(PB_ACCFULL + s * 2 + buf), (G >> 1) & 1);
          tc_fence_after();
          if (blockIdx.x == 0 && threadIdx.x == 0 && gi < 64) g_pair_stamps[gi * 8 + 6] = clock64();
          const uint32_t tacc = tl + (uint32_t)(s * 2 + buf) * 128;
          const bool fwd = g < L;
          const int l = fwd ? g + 1 : 2 * L - g;
          const float wsi = hdr[l - 1];
          uint32_t rn[8];
          tmem_ld8_issue(tacc, rn);
          {
            const int lp = fwd ? (l == L ? L - 1 : 0) : l - 2;   // stash layer the NEXT backward stage reads
            if (lp >= 1) {
              const float4* pf = stash + (size_t)(lp - 1) * 64 * HR + hrow;
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const int col4 = (4 * ch + j) * 8 + w4 * 2;
                asm volatile("prefetch.global.L2 [%0];" ::"l"(pf + (size_t)col4 * HR));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(pf + (size_t)(col4 + 1) * HR));
              }
            }
          }

          if (fwd && l < L) {
            // ---- E_f(l): h_l = sin(w z_l) -> A ; stash c_l = w cos(w z_l) ----
            const float2 sc2 = bc2(wsi * A_SCALE_INV * omega);
            float4* st = stash + (size_t)(l - 1) * 64 * HR + hrow;
            const float4* bw4 = reinterpret_cast<const float4*>(biasw + (l - 1) * H) + w4 * 2;
            float4 bwn0 = __ldg(bw4 + (4 * ch) * 8), bwn1 = __ldg(bw4 + (4 * ch) * 8 + 1);
#pragma unroll 1
            for (int j = 0; j < 4; ++j) {
              const int kb = 4 * ch + j;
              tmem_ld_wait(rn);
              const float2 v[4] = {make_float2(__uint_as_float(rn[0]), __uint_as_float(rn[1])),
                                   make_float2(__uint_as_float(rn[2]), __uint_as_float(rn[3])),
                                   make_float2(__uint_as_float(rn[4]), __uint_as_float(rn[5])),
                                   make_float2(__uint_as_float(rn[6]), __uint_as_float(rn[7]))};
              if (j + 1 < 4) tmem_ld8_issue(tacc + (j + 1) * KB, rn);
              const float2 bb[4] = {make_float2(bwn0.x, bwn0.y), make_float2(bwn0.z, bwn0.w),
                                    make_float2(bwn1.x, bwn1.y), make_float2(bwn1.z, bwn1.w)};
              float2 o[4], cc[4];
#pragma unroll
              for (int p2 = 0; p2 < 4; ++p2) {
                float2 sn, cp;
                uint32_t sx, sy;
                sincos2(__ffma2_rn(v[p2], sc2, bb[p2]), sn, cp, sx, sy);
                o[p2] = __fmul2_rn(sn, bc2(A_SCALE));
                cc[p2] = __fmul2_rn(cp, signed_scale(omega, sx, sy));
              }
              publish(kb, o);
              const int col4 = kb * 8 + w4 * 2;
              st[(size_t)col4 * HR] = make_float4(cc[0].x, cc[0].y, cc[1].x, cc[1].y);
              st[(size_t)(col4 + 1) * HR] = make_float4(cc[2].x, cc[2].y, cc[3].x, cc[3].y);
              if (j + 1 < 4) {
                bwn0 = __ldg(bw4 + (kb + 1) * 8);
                bwn1 = __ldg(bw4 + (kb + 1) * 8 + 1);
              }
            }
          } else if (fwd) {
            // ---- E_f(L): sdf = h_L . w_last + b_last ; A = gl_scale * w_last * c_L ----
            const float2 sc2 = bc2(wsi * A_SCALE_INV * omega);
            const float gls = gl_scale * omega;
            const float4* bw4 = reinterpret_cast<const float4*>(biasw + (l - 1) * H) + w4 * 2;
            float2 acc2 = bc2(0.f);
#pragma unroll 1
            for (int j = 0; j < 4; ++j) {
              const int kb = 4 * ch + j;
              tmem_ld_wait(rn);
              const float2 v[4] = {make_float2(__uint_as_float(rn[0]), __uint_as_float(rn[1])),
                                   make_float2(__uint_as_float(rn[2]), __uint_as_float(rn[3])),
                                   make_float2(__uint_as_float(rn[4]), __uint_as_float(rn[5])),
                                   make_float2(__uint_as_float(rn[6]), __uint_as_float(rn[7]))};
              if (j + 1 < 4) tmem_ld8_issue(tacc + (j + 1) * KB, rn);
              const float4 b0 = __ldg(bw4 + kb * 8), b1 = __ldg(bw4 + kb * 8 + 1);
              const float4 w0 = __ldg(w_last4 + kb * 8), w1 = __ldg(w_last4 + kb * 8 + 1);
              const float2 bb[4] = {make_float2(b0.x, b0.y), make_float2(b0.z, b0.w), make_float2(b1.x, b1.y),
                                    make_float2(b1.z, b1.w)};
              const float2 ww[4] = {make_float2(w0.x, w0.y), make_float2(w0.z, w0.w), make_float2(w1.x, w1.y),
                                    make_float2(w1.z, w1.w)};
              float2 o[4];
#pragma unroll
              for (int p2 = 0; p2 < 4; ++p2) {
                float2 sn, cp;
                uint32_t sx, sy;
                sincos2(__ffma2_rn(v[p2], sc2, bb[p2]), sn, cp, sx, sy);
                acc2 = __ffma2_rn(sn, ww[p2], acc2);
                o[p2] = __fmul2_rn(__fmul2_rn(cp, signed_scale(gls, sx, sy)), ww[p2]);
              }
              publish(kb, o);
            }
            const float acc_sdf = acc2.x + acc2.y;
            rsi[s] = gl_scale_inv;
            if (pi) xch[hrow * 8 + pi] = acc_sdf;
            row_barrier8();
            if (pi == 0) {
              const float* e = xch + hrow * 8;
              sdfr[s] = (((acc_sdf + e[1]) + (e[2] + e[3])) + ((e[4] + e[5]) + (e[6] + e[7]))) + b_last;
              if (valid && sdf_out) sdf_out[grow] = sdfr[s];
            }
            row_barrier8();
          } else if (l > 1) {
            // ---- E_b(l): g_{l-1} = acc / scales ; gp_{l-1} = g_{l-1} * c_{l-1} -> A (row-scaled) ----
            const float sc = wsi * rsi[s];
            float m = 0.f;
            {
              tmem_ld_wait(rn);   // retire the prefetch; re-issued below
              float w32[32];
              tmem_ld32(tacc - 8 * w4 + 32 * w4, w32);   // any thread of the row may scan any columns
#pragma unroll
              for (int i = 0; i < 32; ++i) m = fmaxf(m, fabsf(w32[i]));
              tmem_ld8_issue(tacc, rn);
            }
            xch[hrow * 8 + pi] = m;
            row_barrier8();
            {
              const float* e = xch + hrow * 8;
              m = fmaxf(fmaxf(fmaxf(e[0], e[1]), fmaxf(e[2], e[3])), fmaxf(fmaxf(e[4], e[5]), fmaxf(e[6], e[7])));
            }
            row_barrier8();
            const float new_scale = pow2_scale_for(m * sc * fabsf(omega));
            const float2 scs2 = bc2(sc * new_scale);
            const float4* st = stash + (size_t)(l - 2) * 64 * HR + hrow + (size_t)((4 * ch) * 8 + w4 * 2) * HR;
            float4 c0 = __ldcg(st), c1 = __ldcg(st + HR);
            float4 d0 = __ldcg(st + (size_t)8 * HR), d1 = __ldcg(st + (size_t)9 * HR);
#pragma unroll 2
            for (int j = 0; j < 4; ++j) {
              const int kb = 4 * ch + j;
              tmem_ld_wait(rn);
              float2 o[4];
              o[0] = __fmul2_rn(__fmul2_rn(make_float2(__uint_as_float(rn[0]), __uint_as_float(rn[1])), scs2),
                                make_float2(c0.x, c0.y));
              o[1] = __fmul2_rn(__fmul2_rn(make_float2(__uint_as_float(rn[2]), __uint_as_float(rn[3])), scs2),
                                make_float2(c0.z, c0.w));
              o[2] = __fmul2_rn(__fmul2_rn(make_float2(__uint_as_float(rn[4]), __uint_as_float(rn[5])), scs2),
                                make_float2(c1.x, c1.y));
              o[3] = __fmul2_rn(__fmul2_rn(make_float2(__uint_as_float(rn[6]), __uint_as_float(rn[7])), scs2),
                                make_float2(c1.z, c1.w));
              if (j + 1 < 4) tmem_ld8_issue(tacc + (j + 1) * KB, rn);
              publish(kb, o);
              if ((lane & 7) == 0) {   // the tape lines of this k-block are dead: drop them from L2
                asm volatile("discard.global.L2 [%0], 128;" ::"l"(st + (size_t)(j * 8) * HR) : "memory");
                asm volatile("discard.global.L2 [%0], 128;" ::"l"(st + (size_t)(j * 8 + 1) * HR) : "memory");
              }
              c0 = d0;
              c1 = d1;
              if (j + 2 < 4) {
                d0 = __ldcg(st + (size_t)((j + 2) * 8) * HR);
                d1 = __ldcg(st + (size_t)((j + 2) * 8 + 1) * HR);
              }
            }
            rsi[s] = 1.f / new_scale;
          } else {
            // ---- E_b(1): g_0 = acc / scales ; gp_0 = g_0 * w0 cos(w0 z_0) ; grad = gp_0 W_0 ----
            const float sc = wsi * rsi[s];
            float2 gx2 = bc2(0.f), gy2 = bc2(0.f), gz2 = bc2(0.f);
#pragma unroll 1
            for (int j = 0; j < 4; ++j) {
              const int kb = 4 * ch + j;
              tmem_ld_wait(rn);
              const float2 v[4] = {make_float2(__uint_as_float(rn[0]), __uint_as_float(rn[1])),
                                   make_float2(__uint_as_float(rn[2]), __uint_as_float(rn[3])),
                                   make_float2(__uint_as_float(rn[4]), __uint_as_float(rn[5])),
                                   make_float2(__uint_as_float(rn[6]), __uint_as_float(rn[7]))};
              if (j + 1 < 4) tmem_ld8_issue(tacc + (j + 1) * KB, rn);
#pragma unroll
              for (int p2 = 0; p2 < 4; ++p2) {
                const float4 wa = __ldg(w0p + kb * 32 + p2 * 2), wb = __ldg(w0p + kb * 32 + p2 * 2 + 1);
                const float2 wx = make_float2(wa.x, wa.y), wy = make_float2(wa.z, wa.w), wz = make_float2(wb.x, wb.y);
                const float2 th =
                    __ffma2_rn(wz, pz2, __ffma2_rn(wy, py2, __ffma2_rn(wx, px2, make_float2(wb.z, wb.w))));
                float2 sn, cp;
                uint32_t sx, sy;
                sincos2(th, sn, cp, sx, sy);
                const float2 gp = __fmul2_rn(__fmul2_rn(v[p2], signed_scale(sc, sx, sy)), cp);
                gx2 = __ffma2_rn(gp, wx, gx2);
                gy2 = __ffma2_rn(gp, wy, gy2);
                gz2 = __ffma2_rn(gp, wz, gz2);
              }
            }
            const float gx = gx2.x + gx2.y, gy = gy2.x + gy2.y, gz = gz2.x + gz2.y;
            tc_fence_before();
            // partials of the row's other 7 threads go through this slot's (idle) A half-tile: the 16-byte
            // slot of this row in K-chunk pi - 1.  The barrier in front restates, for racecheck, the ordering
            // the mbarrier chain already gives against the A chunk last stored there.
            unsigned char* gsc = smem + s * PA_SLOT + hrow * 16;
            row_barrier8();
            if (pi) *reinterpret_cast<float4*>(gsc + (pi - 1) * PA_LBO) = make_float4(gx, gy, gz, 0.f);
            row_barrier8();
            float fgx = gx, fgy = gy, fgz = gz;
            if (pi == 0) {
#pragma unroll
              for (int k = 0; k < 7; ++k) {
                const float4 d = *reinterpret_cast<const float4*>(gsc + k * PA_LBO);
                fgx += d.x;
                fgy += d.y;
                fgz += d.z;
              }
              if (valid && grad_out) {
                grad_out[3 * (size_t)grow] = fgx;
                grad_out[3 * (size_t)grow + 1] = fgy;
                grad_out[3 * (size_t)grow + 2] = fgz;
              }
            }
            row_barrier8();
            if (nw.points && pi == 0) {
              // ---- fused Newton step on this row (warps 0 and 1 hold the half-tile's 64 rows) ----
              int* cmp = reinterpret_cast<int*>(smem + PSM_CMP);
              bool still = false;
              int p = 0;
              float nx = 0.f, ny = 0.f, nz = 0.f;
              const float sdf_row = sdfr[s];
              if (valid) {
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
                    const float dn = fmaxf(nrm, 1e-15f);
                    const float len = fminf(nrm, nw.max_step);
                    nx = __fsub_rn(nx, __fmul_rn(__fdiv_rn(mx, dn), len));
                    ny = __fsub_rn(ny, __fmul_rn(__fdiv_rn(my, dn), len));
                    nz = __fsub_rn(nz, __fmul_rn(__fdiv_rn(mz, dn), len));
                    nw.points[3 * (size_t)p] = nx;
                    nw.points[3 * (size_t)p + 1] = ny;
                    nw.points[3 * (size_t)p + 2] = nz;
                  }
                }
              }
              const unsigned bal = __ballot_sync(0xffffffffu, still);
              if (lane == 0) cmp[q & 1] = __popc(bal);
              asm volatile("bar.sync 5, 64;" ::: "memory");
              const int c0 = cmp[0], c1 = cmp[1];
              if (threadIdx.x == 0) cmp[4] = (c0 + c1) ? atomicAdd(nw.count_out, c0 + c1) : 0;
              asm volatile("bar.sync 5, 64;" ::: "memory");
              if (still) {
                const int pos = cmp[4] + ((q & 1) ? c0 : 0) + __popc(bal & ((1u << lane) - 1u));
                nw.act_out[pos] = p;
                if (nw.next_points) {
                  nw.next_points[3 * (size_t)pos] = nx;
                  nw.next_points[3 * (size_t)pos + 1] = ny;
                  nw.next_points[3 * (size_t)pos + 2] = nz;
                }
              }
            }
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // the peer may still be reading this CTA's shared memory / TMEM through the pair MMAs
  if (warp == N_EPI_WARPS + 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

}  // namespace pr

}  // namespace siren
}  // namespace isob200
extern "C" int isob200_siren_pair_stamps(long long* out, int n) {
  using namespace isob200;
  ISO_CUDA(cudaMemcpyFromSymbol(out, siren::pr::g_pair_stamps, sizeof(long long) * (n < 512 ? n : 512)));
  return ISOB200_OK;
}
namespace isob200 {
namespace siren {
// launcher used by siren.cu's C-ABI entry points when the pair kernel is selected
int launch_siren_pair(const float* x, int n_max, const int* n_dev, const void* blob, int n_hidden, float* sdf,
                      float* grad, void* scratch, const Newton& nw, cudaStream_t stream) {
  ISO_CUDA(cudaFuncSetAttribute(pr::siren_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, pr::PSMEM_BYTES));
  const int tiles = (n_max + TM - 1) / TM;
  const int pairs = tiles < kNumSMs / 2 ? tiles : kNumSMs / 2;
  pr::siren_pair_kernel<<<2 * pairs, THREADS, pr::PSMEM_BYTES, stream>>>(
      x, n_max, n_dev, (const unsigned char*)blob, n_hidden, sdf, grad, (float*)scratch, nw);
  ISO_CHECK_LAUNCH("siren_pair_kernel");
  return ISOB200_OK;
}

}  // namespace siren
}  // namespace isob200
