// DSS elliptical-splat rasteriser: backward, per-point visibility and the fused RGBA blend, sm_100a.
//
// Replaces
//   DSS/csrc/rasterize_points_backward.cu:30-212  RasterizePointsBackwardCudaFastKernel
//   DSS/csrc/rasterize_points.cu:673-760          RasterizePointsOccBackwardCudaKernel (slow path)
//   DSS/csrc/rasterize_points.cu:823-846          ZbufBackwardKernel
//   DSS/utils/__init__.py:378-399                 get_per_point_visibility_mask (unique + scatter)
//   DSS/core/renderer.py:53-78                    exp(-Q/2)*scaler weights + pytorch3d
//                                                 NormWeightedCompositor + alpha concat
//
// Occupancy backward -- B200 design.  The reference is PIXEL-centric: one thread per pixel with
// grad_occ != 0 walks a 2-D FRNN grid of the visible points (built per call with insert / scan /
// counting-sort through frnn._C and a python loop) and issues two fp32 atomics per (pixel, point)
// pair -- ~800 pairs per pixel at the default radii_backward_scaler = 10, nondeterministic sums.
// The pair relation is symmetric, so here it is POINT-centric: one WARP per visible point sweeps
// the (2r+1)^2 pixel window of grad_occ around the point with coalesced row reads (the 1 MB
// grad_occ image of a view stays L2 resident), accumulates in registers, and finishes with a
// warp-shuffle reduction and ONE plain store per coordinate: no grid build, no atomics,
// run-to-run deterministic.  The predicate per (pixel, point) pair is the reference's, in its
// fp32 form (dist2 = fma(dx, dx, dy*dy) <= r^2, ...), so the set of contributing pairs is equal.
// The fast path (mode 0) goes one step further (splat_occ_backward_tiled_kernel): points are binned by the
// 16x16 tile of their centre, one CTA per tile stages the neighbourhood's non-zero gradient pixels in shared
// memory once and every THREAD owns a point -- 0.77 -> 0.28 ms at BASELINE config 4; the warp-per-point
// kernels remain for the slow-path semantics (mode 1) and as the plain reference sweep (no workspace).
// NOT reproduced: the reference closes the last 2-D grid cell of views n >= 1 with a local count
// while its offsets are packed-global (rasterize_points_backward.cu:124-126), silently dropping
// that cell's points; here every in-radius pair contributes.
#include "common.cuh"
#include "scan.cuh"
#include <float.h>

namespace isob200 {

__device__ __forceinline__ float pix_to_ndc_b(int i, float fS) {
  return __fadd_rn(__fdiv_rn(__fadd_rn((float)(2 * i), 1.0f), fS), -1.0f);
}

// pixel indices i in [0,S) that can satisfy |ndc(i) - p| <= r  (superset, lo > hi when empty)
__device__ __forceinline__ void pixel_window(float p, float r, int S, float fS, int& lo, int& hi) {
  const float e_lo = (p - r + 1.0f) * fS * 0.5f - 0.5f;
  const float e_hi = (p + r + 1.0f) * fS * 0.5f - 0.5f;
  if (!(e_lo <= (float)S) || !(e_hi >= -1.0f)) { lo = 1; hi = 0; return; }
  lo = (int)fmaxf(ceilf(e_lo) - 1.0f, 0.0f);
  hi = (int)fminf(floorf(e_hi) + 1.0f, (float)(S - 1));
}

// rasterization_utils.cuh:38-44 (device eps_denom): sign(0) = 0, unlike the python helper
__device__ __forceinline__ float eps_denom_dev(float d, float eps) {
  const float s = (float)((0.0f < d) - (d < 0.0f));
  return __fmul_rn(s, fmaxf(fabsf(d), eps));
}

// mode 0: fast-path semantics (rasterize_points_backward.cu): radius rs[n], circle test.
// mode 1: slow-path semantics (rasterize_points.cu:673-760): per-point box radii * radii_s.
template <int MODE>
__global__ void __launch_bounds__(256)
splat_occ_backward_kernel(const float* __restrict__ points, const float* __restrict__ radii,
                          const unsigned char* __restrict__ visible,
                          const int64_t* __restrict__ first_idx, const int64_t* __restrict__ num_points,
                          const float* __restrict__ rs, float radii_s, const float* __restrict__ grad_occ,
                          int H, int W, float* __restrict__ grad_out, int out_stride) {
  const int n = blockIdx.y;
  const long long first = first_idx[n], num = num_points[n];
  const int lane = threadIdx.x & 31;
  const long long warps = (long long)gridDim.x * (blockDim.x >> 5);
  const float fW = (float)W, fH = (float)H;
  const float* g_img = grad_occ + (size_t)n * H * W;
  for (long long i = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); i < num; i += warps) {
    const long long p = first + i;
    float gx = 0.f, gy = 0.f;
    const float px = points[3 * p], py = points[3 * p + 1], pz = points[3 * p + 2];
    const bool live = (visible == nullptr || visible[p]) && !(pz < 0.f || fabsf(py) > 1.0f || fabsf(px) > 1.0f);
    if (live) {
      const float rx = radii[2 * p], ry = radii[2 * p + 1];
      float wx, wy, r2 = 0.f, bx = rx, by = ry;
      if (MODE == 0) {
        const float r = rs[n];
        r2 = __fmul_rn(r, r);
        wx = wy = r;
      } else {
        wx = __fmul_rn(rx, radii_s);           // radiix (:724-725)
        wy = __fmul_rn(ry, radii_s);
        bx = __fdiv_rn(wx, radii_s);           // radiix / radii_s (:741)
        by = __fdiv_rn(wy, radii_s);
      }
      int xl, xh, yl, yh;
      pixel_window(px, wx, W, fW, xl, xh);
      pixel_window(py, wy, H, fH, yl, yh);
      if (xl <= xh && yl <= yh) {
        const int c_lo = W - 1 - xh, c_hi = W - 1 - xl;   // output columns (x is flipped)
        // lane = column: dx is row independent, so it is formed once per 32-column chunk; the
        // per-row dy comes from a lane that computed ndc(y) for 32 rows at once (one shuffle per
        // row instead of an int->float + IEEE division per pixel).
        for (int cb = c_lo; cb <= c_hi; cb += 32) {
          const int col = cb + lane;
          const bool col_ok = col <= c_hi;
          const float dx = __fsub_rn(pix_to_ndc_b(W - 1 - col, fW), px);
          const bool x_out = fabsf(dx) > bx;              // outside the splat's radii box in x
          const bool x_far = (MODE == 1) && (fabsf(dx) > wx);
          const float* g_col = g_img + (col_ok ? col : c_lo);
          for (int rb = yl; rb <= yh; rb += 32) {
            const float dy_l = __fsub_rn(pix_to_ndc_b(min(rb + lane, H - 1), fH), py);
            const int nr = min(32, yh - rb + 1);
#pragma unroll 4
            for (int r = 0; r < nr; ++r) {
              const float g = g_col[(size_t)(H - 1 - (rb + r)) * W];
              const float dy = __shfl_sync(0xffffffffu, dy_l, r);
              if (g == 0.0f || !col_ok) continue;
              const float d2 = __fmaf_rn(dx, dx, __fmul_rn(dy, dy));   // SASS: FMUL dy*dy ; FFMA dx
              if (MODE == 0) {
                if (d2 > r2) continue;
              } else {
                if (x_far || fabsf(dy) > wy) continue;
              }
              if (g > 0.0f && (x_out || fabsf(dy) > by)) continue;
              // dx / eps_denom(d2, 1e-10) * g with the device eps_denom's sign(0) = 0 (0/0 -> NaN kept);
              // the quotient goes through the SFU reciprocal (1 ulp): gradients are fp32 sums whose
              // order the reference itself does not fix
              const float den = (d2 > 0.0f) ? fmaxf(d2, 1e-10f) : 0.0f;
              float inv;
              asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv) : "f"(den));
              const float w = inv * g;
              gx = fmaf(dx, w, gx);
              gy = fmaf(dy, w, gy);
            }
          }
        }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      gx += __shfl_xor_sync(0xffffffffu, gx, o);
      gy += __shfl_xor_sync(0xffffffffu, gy, o);
    }
    if (lane == 0) {
      grad_out[(size_t)p * out_stride + 0] = gx;
      grad_out[(size_t)p * out_stride + 1] = gy;
    }
  }
}

// ---- sparse form of grad_occ: per 16x16 output tile, the list of pixels with a non-zero gradient ----
// Occupancy gradients are zero wherever the rendered and target masks agree, i.e. almost everywhere
// except a band around silhouettes; the point-centric sweep above still visits every pixel of its
// window.  Two tiny kernels (+ the scan) turn the image into per-tile record lists {xf, yf, g, -} in
// pixel order (deterministic), and the hybrid kernel below walks, for every tile its window touches,
// either the tile's list (sparse tile) or the pixels themselves (dense tile).
constexpr int GT = 16;   // tile side of the gradient-pixel lists

__global__ void __launch_bounds__(256)
gradpix_count_kernel(const float* __restrict__ grad_occ, int H, int W, int TX, int TY, int* __restrict__ tile_cnt,
                     int* __restrict__ ticket, int* __restrict__ tile_off) {
  const int t = blockIdx.x;                       // n*TY*TX + ty*TX + tx
  const int n = t / (TX * TY), tr = t - n * TX * TY;
  const int ty = tr / TX, tx = tr - ty * TX;
  const int col = tx * GT + (threadIdx.x & 15), row = ty * GT + (threadIdx.x >> 4);
  const bool nz = col < W && row < H && grad_occ[((size_t)n * H + row) * W + col] != 0.0f;
  const int c = __syncthreads_count(nz);
  if (threadIdx.x == 0) tile_cnt[t] = c;
  if (ticket) last_cta_exclusive_scan(ticket, (int)gridDim.x, tile_cnt, tile_off, (int)gridDim.x, nullptr, nullptr);
}

__global__ void __launch_bounds__(256)
gradpix_fill_kernel(const float* __restrict__ grad_occ, int H, int W, int TX, int TY,
                    const int* __restrict__ tile_off, float4* __restrict__ recs) {
  __shared__ int warp_cnt[8];
  const int t = blockIdx.x;
  const int n = t / (TX * TY), tr = t - n * TX * TY;
  const int ty = tr / TX, tx = tr - ty * TX;
  const int col = tx * GT + (threadIdx.x & 15), row = ty * GT + (threadIdx.x >> 4);
  float g = 0.0f;
  if (col < W && row < H) g = grad_occ[((size_t)n * H + row) * W + col];
  const bool nz = g != 0.0f;
  const unsigned m = __ballot_sync(0xffffffffu, nz);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) warp_cnt[w] = __popc(m);
  __syncthreads();
  int base = 0;
  for (int i = 0; i < w; ++i) base += warp_cnt[i];
  if (nz) {
    const int slot = tile_off[t] + base + __popc(m & ((1u << lane) - 1u));
    recs[slot] = make_float4(pix_to_ndc_b(W - 1 - col, (float)W), pix_to_ndc_b(H - 1 - row, (float)H), g, 0.0f);
  }
}

// one (pixel, point) pair of the occupancy backward; returns false when the pair does not contribute
template <int MODE>
__device__ __forceinline__ void occ_pair(float dx, float dy, float g, float r2, float wx, float wy, float bx,
                                         float by, float& gx, float& gy) {
  const float d2 = __fmaf_rn(dx, dx, __fmul_rn(dy, dy));   // SASS of the reference: FMUL dy*dy ; FFMA dx
  if (MODE == 0) {
    if (d2 > r2) return;
  } else {
    if (fabsf(dx) > wx || fabsf(dy) > wy) return;
  }
  if (g > 0.0f && (fabsf(dx) > bx || fabsf(dy) > by)) return;
  const float den = (d2 > 0.0f) ? fmaxf(d2, 1e-10f) : 0.0f;
  float inv;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv) : "f"(den));
  const float w = inv * g;
  gx = fmaf(dx, w, gx);
  gy = fmaf(dy, w, gy);
}

template <int MODE>
__global__ void __launch_bounds__(256)
splat_occ_backward_hybrid_kernel(const float* __restrict__ points, const float* __restrict__ radii,
                                 const unsigned char* __restrict__ visible,
                                 const int64_t* __restrict__ first_idx, const int64_t* __restrict__ num_points,
                                 const float* __restrict__ rs, float radii_s, const float* __restrict__ grad_occ,
                                 const int* __restrict__ tile_cnt, const int* __restrict__ tile_off,
                                 const float4* __restrict__ recs, int H, int W, int TX, int TY,
                                 float* __restrict__ grad_out, int out_stride) {
  const int n = blockIdx.y;
  const long long first = first_idx[n], num = num_points[n];
  const int lane = threadIdx.x & 31;
  const long long warps = (long long)gridDim.x * (blockDim.x >> 5);
  const float fW = (float)W, fH = (float)H;
  const float* g_img = grad_occ + (size_t)n * H * W;
  const int* cnt_n = tile_cnt + (size_t)n * TX * TY;
  const int* off_n = tile_off + (size_t)n * TX * TY;
  for (long long i = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); i < num; i += warps) {
    const long long p = first + i;
    float gx = 0.f, gy = 0.f;
    const float px = points[3 * p], py = points[3 * p + 1], pz = points[3 * p + 2];
    const bool live = (visible == nullptr || visible[p]) && !(pz < 0.f || fabsf(py) > 1.0f || fabsf(px) > 1.0f);
    if (live) {
      const float rx = radii[2 * p], ry = radii[2 * p + 1];
      float wx, wy, r2 = 0.f, bx = rx, by = ry;
      if (MODE == 0) {
        const float r = rs[n];
        r2 = __fmul_rn(r, r);
        wx = wy = r;
      } else {
        wx = __fmul_rn(rx, radii_s);
        wy = __fmul_rn(ry, radii_s);
        bx = __fdiv_rn(wx, radii_s);
        by = __fdiv_rn(wy, radii_s);
      }
      int xl, xh, yl, yh;
      pixel_window(px, wx, W, fW, xl, xh);
      pixel_window(py, wy, H, fH, yl, yh);
      if (xl <= xh && yl <= yh) {
        const int c_lo = W - 1 - xh, c_hi = W - 1 - xl;   // window in output coordinates
        const int r_lo = H - 1 - yh, r_hi = H - 1 - yl;
        for (int ty = r_lo / GT; ty <= r_hi / GT; ++ty) {
          for (int tx = c_lo / GT; tx <= c_hi / GT; ++tx) {
            const int t = ty * TX + tx;
            const int cnt = cnt_n[t];
            if (cnt == 0) continue;
            if (cnt <= 96) {
              // sparse tile: walk its non-zero pixels, 32 records per step
              const float4* rl = recs + off_n[t];
              for (int k = lane; k < cnt; k += 32) {
                const float4 rc = rl[k];
                occ_pair<MODE>(__fsub_rn(rc.x, px), __fsub_rn(rc.y, py), rc.z, r2, wx, wy, bx, by, gx, gy);
              }
            } else {
              // dense tile: sweep the part of the window inside it (lane = column)
              const int ca = max(c_lo, tx * GT), cb = min(c_hi, tx * GT + GT - 1);
              const int ra = max(r_lo, ty * GT), rb = min(r_hi, ty * GT + GT - 1);
              const int ncol = cb - ca + 1;                     // <= 16: two rows per warp step
              const int col = ca + (lane & 15);
              const bool col_ok = (lane & 15) < ncol;
              const float dx = __fsub_rn(pix_to_ndc_b(W - 1 - col, fW), px);
              for (int row = ra + (lane >> 4); row <= rb; row += 2) {
                if (!col_ok) continue;
                const float g = g_img[(size_t)row * W + col];
                if (g == 0.0f) continue;
                const float dy = __fsub_rn(pix_to_ndc_b(H - 1 - row, fH), py);
                occ_pair<MODE>(dx, dy, g, r2, wx, wy, bx, by, gx, gy);
              }
            }
          }
        }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      gx += __shfl_xor_sync(0xffffffffu, gx, o);
      gy += __shfl_xor_sync(0xffffffffu, gy, o);
    }
    if (lane == 0) {
      grad_out[(size_t)p * out_stride + 0] = gx;
      grad_out[(size_t)p * out_stride + 1] = gy;
    }
  }
}

// ---- tiled form of the fast-path occupancy backward (mode 0) -----------------------------------------------
// The warp-per-point walk above spends most of its instructions on bookkeeping: per point and touched tile a
// count + offset load, a loop set-up and one 32-record step that is three-quarters empty (a 16x16 tile holds ~26
// non-zero pixels when 10 % of the image carries a gradient), ~360 warp instructions per point.  Points that fall
// into the same 16x16 pixel tile see the same (2R+1)^2 neighbourhood of gradient tiles, so here the live points are
// binned by the tile of their centre (count / scan / fill, like the gradient pixels), one CTA per non-empty
// (view, tile) stages the neighbourhood's gradient records in shared memory ONCE and every thread owns one point:
// the inner loop is a broadcast LDS.128 of a record + the reference's fp32 pair predicate, 32 different points per
// warp instruction and no shuffles.  Sums are per point in a fixed order (tile order, then pixel order): run-to-run
// deterministic, no atomics on the result.
constexpr int OCC_CHUNK = 2304;   // staged records per pass: 9 full tiles, 36 KB of shared memory

// pixel cell of an NDC coordinate (cell i spans [2i/S - 1, (2i+2)/S - 1]), clamped to the image
__device__ __forceinline__ int ndc_to_cell(float p, int S) {
  const int c = __float2int_rd((p + 1.0f) * (float)S * 0.5f);
  return min(max(c, 0), S - 1);
}

__device__ __forceinline__ bool occ_point_live(const float* __restrict__ points, const unsigned char* __restrict__ visible,
                                               long long p) {
  const float px = points[3 * p], py = points[3 * p + 1], pz = points[3 * p + 2];
  return (visible == nullptr || visible[p]) && !(pz < 0.f || fabsf(py) > 1.0f || fabsf(px) > 1.0f);
}

// FILL = false: count the live points per (view, output tile) and zero the gradient rows of the others;
// FILL = true : write the point ids into the tile's slice (cursor = running count per tile)
template <bool FILL>
__global__ void __launch_bounds__(256)
occ_point_bin_kernel(const float* __restrict__ points, const unsigned char* __restrict__ visible,
                     const int64_t* __restrict__ first_idx, const int64_t* __restrict__ num_points, int H, int W,
                     int TX, int TY, int* __restrict__ pcnt, const int* __restrict__ poff, int* __restrict__ plist,
                     float* __restrict__ grad_out, int out_stride) {
  const int n = blockIdx.y;
  const long long first = first_idx[n], num = num_points[n];
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < num;
       i += (long long)gridDim.x * blockDim.x) {
    const long long p = first + i;
    if (!occ_point_live(points, visible, p)) {
      if (!FILL) {
        grad_out[(size_t)p * out_stride + 0] = 0.f;
        grad_out[(size_t)p * out_stride + 1] = 0.f;
      }
      continue;
    }
    // output coordinates are flipped in both axes (column W-1-ix shows NDC cell ix)
    const int col = W - 1 - ndc_to_cell(points[3 * p], W), row = H - 1 - ndc_to_cell(points[3 * p + 1], H);
    const int t = (n * TY + row / GT) * TX + col / GT;
    if (FILL) plist[poff[t] + atomicAdd(pcnt + t, 1)] = (int)i;
    else atomicAdd(pcnt + t, 1);
  }
}

// Count form of the kernel above with the privatisation of splat_tile_count_priv_kernel (splat.cu): a CTA takes a
// contiguous chunk of one view's points, counts into a shared-memory histogram of that view's tiles, adds the non-zero
// bins to the global counters, and the last CTA turns the counts into offsets (no scan launches).
constexpr int OCC_PRIV_MAX_TILES = 4096;
constexpr int OCC_PRIV_CHUNK = 4096;

__global__ void __launch_bounds__(256)
occ_point_count_priv_kernel(const float* __restrict__ points, const unsigned char* __restrict__ visible,
                            const int64_t* __restrict__ first_idx, const int64_t* __restrict__ num_points, int H, int W,
                            int TX, int TY, int* __restrict__ pcnt, int* __restrict__ ticket, int* __restrict__ poff,
                            float* __restrict__ grad_out, int out_stride) {
  extern __shared__ int s_hist[];   // [TX*TY]
  const int n = blockIdx.y, nt = TX * TY;
  const long long first = first_idx[n], num = num_points[n];
  const long long chunk = (num + gridDim.x - 1) / gridDim.x;
  const long long b = (long long)blockIdx.x * chunk, e = min(num, b + chunk);
  if (b < e) {
    for (int i = threadIdx.x; i < nt; i += blockDim.x) s_hist[i] = 0;
    __syncthreads();
    for (long long i0 = b + threadIdx.x; i0 < e; i0 += 4 * blockDim.x) {
      float px[4], py[4], pz[4];
      bool vis[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {                 // four points' loads in flight
        const long long i = i0 + (long long)u * blockDim.x;
        const long long p = first + (i < e ? i : i0);
        px[u] = points[3 * p]; py[u] = points[3 * p + 1]; pz[u] = points[3 * p + 2];
        vis[u] = visible == nullptr || visible[p];
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const long long i = i0 + (long long)u * blockDim.x;
        if (i >= e) break;
        const long long p = first + i;
        if (!(vis[u] && !(pz[u] < 0.f || fabsf(py[u]) > 1.0f || fabsf(px[u]) > 1.0f))) {   // occ_point_live
          grad_out[(size_t)p * out_stride + 0] = 0.f;
          grad_out[(size_t)p * out_stride + 1] = 0.f;
          continue;
        }
        const int col = W - 1 - ndc_to_cell(px[u], W), row = H - 1 - ndc_to_cell(py[u], H);
        atomicAdd(&s_hist[(row / GT) * TX + col / GT], 1);
      }
    }
    __syncthreads();
    int* cnt = pcnt + (size_t)n * nt;
    for (int i = threadIdx.x; i < nt; i += blockDim.x)
      if (s_hist[i]) atomicAdd(&cnt[i], s_hist[i]);
  }
  last_cta_exclusive_scan(ticket, (int)(gridDim.x * gridDim.y), pcnt, poff, nt * (int)gridDim.y, nullptr, nullptr);
}

__global__ void __launch_bounds__(256)
splat_occ_backward_tiled_kernel(const float* __restrict__ points, const float* __restrict__ radii,
                                const int64_t* __restrict__ first_idx, const float* __restrict__ rs,
                                const int* __restrict__ tile_cnt, const int* __restrict__ tile_off,
                                const float4* __restrict__ recs, const int* __restrict__ pcnt,
                                const int* __restrict__ poff, const int* __restrict__ plist, int H, int W, int TX,
                                int TY, float* __restrict__ grad_out, int out_stride) {
  __shared__ float4 srec[OCC_CHUNK];
  __shared__ int s_cnt[64], s_off[64], s_pre[65], s_take;
  const int n = blockIdx.y, tr = blockIdx.x;
  const int t = n * TX * TY + tr;
  const int np = pcnt[t];
  if (np == 0) return;
  const int ty = tr / TX, tx = tr - ty * TX;
  const long long first = first_idx[n];
  const int* pl = plist + poff[t];
  const float r = rs[n];
  const float r2 = __fmul_rn(r, r);
  // tiles a point of this tile can reach: its centre is at most half a pixel from a pixel centre of the tile, a
  // contributing pixel at most r from the centre (one pixel of slack for the fp32 rounding of both)
  const float rpx = r * (float)W * 0.5f + 1.0f, rpy = r * (float)H * 0.5f + 1.0f;
  const int tx0 = max(0, __float2int_rd(((float)(tx * GT) - rpx) / GT));
  const int tx1 = min(TX - 1, __float2int_rd(((float)(tx * GT + GT - 1) + rpx) / GT));
  const int ty0 = max(0, __float2int_rd(((float)(ty * GT) - rpy) / GT));
  const int ty1 = min(TY - 1, __float2int_rd(((float)(ty * GT + GT - 1) + rpy) / GT));
  const int nbx = tx1 - tx0 + 1, nb = nbx * (ty1 - ty0 + 1);
  const int* cnt_n = tile_cnt + (size_t)n * TX * TY;
  const int* off_n = tile_off + (size_t)n * TX * TY;
  bool first_pass = true;
  for (int b0 = 0; b0 < nb;) {
    // ---- stage the records of as many neighbour tiles as fit (each holds <= 256) ----
    __syncthreads();   // previous pass has finished reading srec / s_*
    if (threadIdx.x < 64) {
      const int b = b0 + threadIdx.x;
      int c = 0, o = 0;
      if (b < nb) {
        const int tt = (ty0 + b / nbx) * TX + tx0 + b % nbx;
        c = cnt_n[tt];
        o = off_n[tt];
      }
      s_cnt[threadIdx.x] = c;
      s_off[threadIdx.x] = o;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      int acc = 0, k = 0;
      for (; k < 64 && b0 + k < nb && acc + s_cnt[k] <= OCC_CHUNK; ++k) {
        s_pre[k] = acc;
        acc += s_cnt[k];
      }
      s_pre[k] = acc;   // k <= 64: total records of the tiles taken
      s_take = k;       // tiles taken in this pass (>= 1: a tile never exceeds the chunk)
    }
    __syncthreads();
    const int ntake = s_take;
    const int nrec = s_pre[ntake];
    for (int k = 0; k < ntake; ++k) {
      const float4* src = recs + s_off[k];
      float4* dst = srec + s_pre[k];
      for (int j = threadIdx.x; j < s_cnt[k]; j += blockDim.x) dst[j] = src[j];
    }
    __syncthreads();
    // ---- one point per thread against every staged record ----
    if (nrec > 0 || first_pass) {
      for (int i = threadIdx.x; i < np; i += blockDim.x) {
        const long long p = first + pl[i];
        const float px = points[3 * p], py = points[3 * p + 1];
        const float bx = radii[2 * p], by = radii[2 * p + 1];
        float gx = 0.f, gy = 0.f;
        for (int k = 0; k < nrec; ++k) {
          const float4 rc = srec[k];
          occ_pair<0>(__fsub_rn(rc.x, px), __fsub_rn(rc.y, py), rc.z, r2, r, r, bx, by, gx, gy);
        }
        float* g = grad_out + (size_t)p * out_stride;
        if (first_pass) { g[0] = gx; g[1] = gy; }
        else { g[0] += gx; g[1] += gy; }
      }
    }
    first_pass = false;
    b0 += ntake;
  }
}

// ---- per-view search radius: median(radii of the view's visible points) * radii_s -------------
// (rasterizer.py:884; torch.median = lower middle of the flattened (n_visible, 2) values).  Exact
// radix select on the order-preserving uint image of the floats: 4 passes of 8 bits, one
// histogram sweep + one tiny select kernel per pass; no sort, no host round trip.
__device__ __forceinline__ unsigned f2ord_b(float f) {
  const unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f_b(unsigned u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

// state per view: [0] prefix bits fixed so far, [1] rank still to descend, [2] done flag, [3] blocks finished
// One launch per 8-bit digit: every block histograms its share into shared memory and adds it to the view's 256
// global bins; the LAST block of a view to finish (ticket on state[3]) picks the digit, narrows the prefix and
// clears the bins for the next pass -- no separate select launch between the passes.
__global__ void __launch_bounds__(256)
median_pass_kernel(const float* __restrict__ radii, const unsigned char* __restrict__ visible,
                   const int64_t* __restrict__ first_idx, const int64_t* __restrict__ num_points, int shift,
                   int first_pass, int last_pass, float radii_s, unsigned* __restrict__ state,
                   unsigned* __restrict__ hist, float* __restrict__ rs) {
  __shared__ unsigned sh[256];
  __shared__ bool s_last;
  const int n = blockIdx.y;
  sh[threadIdx.x] = 0;
  __syncthreads();
  const long long first = first_idx[n], num = num_points[n];
  const unsigned prefix = state[n * 4 + 0];
  const unsigned himask = (shift >= 24) ? 0u : (0xffffffffu << (shift + 8));
  // a thread's consecutive samples mostly share the digit in the high passes (same exponent): count runs in a
  // register and touch the shared bin once per run instead of once per sample.  Four points (one float2 of radii +
  // one visibility byte each) are loaded before any is used: with one dependent load per iteration the pass ran
  // at the memory latency (~25 us for 21 MB), not the bandwidth.
  unsigned cur = 0, run = 0;
  const float2* r2 = reinterpret_cast<const float2*>(radii) + first;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; i0 < num; i0 += 4 * stride) {
    float2 rv[4];
    bool ok[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long long i = i0 + u * stride;
      ok[u] = i < num;
      rv[u] = ok[u] ? __ldg(r2 + i) : make_float2(0.f, 0.f);
      if (ok[u] && visible) ok[u] = visible[first + i] != 0;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (!ok[u]) continue;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const unsigned w = f2ord_b(c ? rv[u].y : rv[u].x);
        if (((w ^ prefix) & himask) != 0) continue;
        const unsigned dig = (w >> shift) & 255u;
        if (dig == cur) { ++run; continue; }
        if (run) atomicAdd(&sh[cur], run);
        cur = dig;
        run = 1;
      }
    }
  }
  if (run) atomicAdd(&sh[cur], run);
  __syncthreads();
  if (sh[threadIdx.x]) atomicAdd(&hist[n * 256 + threadIdx.x], sh[threadIdx.x]);
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(&state[n * 4 + 3], 1u) == gridDim.x - 1;
  __syncthreads();
  if (!s_last) return;
  // ---- last block of this view: all 256 bins are final ----
  unsigned* h = hist + n * 256;
  sh[threadIdx.x] = __ldcg(h + threadIdx.x);   // the other blocks' atomics live in L2
  h[threadIdx.x] = 0;                          // ready for the next pass
  __syncthreads();
  if (threadIdx.x >= 32) return;
  // warp 0 picks the digit: lane l owns bins [8l, 8l + 8); a shuffle scan over the lane totals finds the lane that
  // holds rank k, that lane walks its eight bins (a single thread walking all 256 was ~8 us per pass on the
  // critical path of a 4-pass chain)
  unsigned* s = state + n * 4;
  const int lane = threadIdx.x;
  unsigned v[8], tot = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) { v[j] = sh[lane * 8 + j]; tot += v[j]; }
  unsigned inc = tot;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  const unsigned all = __shfl_sync(0xffffffffu, inc, 31);
  unsigned done = s[2], k = s[1];
  const unsigned pre = s[0];
  __syncwarp();
  if (first_pass) {
    done = (all == 0);
    k = all ? (all - 1) / 2 : 0u;
  }
  const unsigned exc = inc - tot;
  const bool mine = !done && (k < all ? (k >= exc && k < inc) : lane == 31);
  if (mine) {
    unsigned acc = exc;
    int j = 0;
    for (; j < 7; ++j) {
      if (acc + v[j] > k) break;
      acc += v[j];
    }
    s[0] = pre | (((unsigned)(lane * 8 + j)) << shift);
    s[1] = k - acc;
    if (last_pass) rs[n] = __fmul_rn(ord2f_b(s[0]), radii_s);
  }
  if (lane == 0) {
    s[3] = 0;
    if (first_pass) {
      s[2] = done;
      if (done) s[1] = 0;
    }
    if (last_pass && done) rs[n] = 0.0f;
  }
}

// z_grad[idx] += grad_zbuf; zero gradients skipped, stop at the first -1 (rasterize_points.cu:835-843)
__global__ void __launch_bounds__(256)
splat_zbuf_backward_kernel(const int* __restrict__ idx, const float* __restrict__ grad_zbuf,
                           long long npix, int K, float* __restrict__ z_grad, int stride) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < npix;
       i += (long long)gridDim.x * blockDim.x) {
    for (int k = 0; k < K; ++k) {
      const float g = grad_zbuf[i * K + k];
      if (g == 0.0f) continue;
      const int p = idx[i * K + k];
      if (p < 0) break;
      atomicAdd(z_grad + (size_t)p * stride, g);
    }
  }
}

// visible[p] = 1 for every id in any of the K slots of an "active" pixel:
//   mask == nullptr: active <=> idx[pixel, 0] >= 0   (EllipticalRasterizer.backward, rasterizer.py:851-857)
//   mask != nullptr: active <=> mask[pixel] != 0     (occupancy: get_per_point_visibility_mask)
__global__ void __launch_bounds__(256)
splat_visibility_kernel(const int* __restrict__ idx, const float* __restrict__ mask, long long npix, int K,
                        long long P, unsigned char* __restrict__ visible) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < npix * K;
       i += (long long)gridDim.x * blockDim.x) {
    const long long pix = i / K;
    const bool active = mask ? (mask[pix] != 0.0f) : (idx[pix * K] >= 0);
    const int p = idx[i];
    if (active && p >= 0 && p < P) visible[p] = 1;
  }
}

// RGBA = [ sum_k w_k f[idx_k] / max(sum_k w_k, eps) , occ ],  w_k = exp(-0.5 q_k) * scaler[idx_k]
// over slots with idx >= 0 (renderer.py:53-78; the normalisation is pytorch3d's norm_weighted_sum).
// Optionally stores the weights (N,S,S,K) for the backward.
template <int C>
__global__ void __launch_bounds__(256)
splat_blend_kernel(const int* __restrict__ idx, const float* __restrict__ qvalue,
                   const float* __restrict__ occ, const float* __restrict__ scaler,
                   const float* __restrict__ feat, int feat_stride, long long npix, int K, float eps,
                   float* __restrict__ out, float* __restrict__ weights_out) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < npix;
       i += (long long)gridDim.x * blockDim.x) {
    float acc[C];
#pragma unroll
    for (int c = 0; c < C; ++c) acc[c] = 0.f;
    float sw = 0.f;
    for (int k = 0; k < K; ++k) {
      const int p = idx[i * K + k];
      float w = 0.f;
      if (p >= 0) {
        w = expf(-0.5f * qvalue[i * K + k]) * (scaler ? scaler[p] : 1.0f);
        sw += w;
#pragma unroll
        for (int c = 0; c < C; ++c) acc[c] += w * feat[(size_t)p * feat_stride + c];
      }
      if (weights_out) weights_out[i * K + k] = w;
    }
    const float den = fmaxf(sw, eps);
#pragma unroll
    for (int c = 0; c < C; ++c) out[i * (C + 1) + c] = acc[c] / den;
    out[i * (C + 1) + C] = occ[i];
  }
}

// d RGBA / d feat: grad_feat[idx_k, c] += w_k / max(sum w, eps) * grad_out[c]
template <int C>
__global__ void __launch_bounds__(256)
splat_blend_backward_kernel(const int* __restrict__ idx, const float* __restrict__ weights,
                            const float* __restrict__ grad_out, long long npix, int K, float eps,
                            float* __restrict__ grad_feat, int feat_stride) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < npix;
       i += (long long)gridDim.x * blockDim.x) {
    float sw = 0.f;
    for (int k = 0; k < K; ++k) sw += (idx[i * K + k] >= 0) ? weights[i * K + k] : 0.f;
    const float inv = 1.0f / fmaxf(sw, eps);
    float g[C];
#pragma unroll
    for (int c = 0; c < C; ++c) g[c] = grad_out[i * (C + 1) + c];
    for (int k = 0; k < K; ++k) {
      const int p = idx[i * K + k];
      if (p < 0) continue;
      const float a = weights[i * K + k] * inv;
#pragma unroll
      for (int c = 0; c < C; ++c) atomicAdd(&grad_feat[(size_t)p * feat_stride + c], a * g[c]);
    }
  }
}

}  // namespace isob200

using namespace isob200;

extern "C" {

// == DSS._C._splat_points_occ_fast_cuda_backward + the python that feeds it
//    (rasterizer.py:850-966; rasterize_points_backward.cu:227-322) when mode == 0, and
// == DSS._C._splat_points_occ_backward (rasterize_points.h:341-386) when mode == 1.
//   visible : (P,) uint8 from isob200_splat_visibility, or NULL = every point participates
//   rs      : (N,) per-view search radius (mode 0);  radii_s: box scale (mode 1)
//   grad_out: rows of `out_stride` floats; columns 0,1 of EVERY row are written (0 when the point
//             does not participate), so the caller needs no zero fill.
size_t isob200_splat_occ_backward_ws_bytes(int N, int H, int W, long long total_points) {
  const size_t nt = (size_t)N * div_up(W, GT) * div_up(H, GT);
  // gradient-pixel tiles: counts, offsets, scan scratch, records; point bins: counts, offsets, ids
  return align_up(nt * 4) * 4 + align_up(scan_ws_bytes((int)nt, 1)) + align_up((size_t)N * H * W * sizeof(float4)) +
         align_up((size_t)(total_points > 0 ? total_points : 0) * 4);
}

//   total_points: rows of points / radii / grad_out (packed over the N clouds)
//   ws / ws_bytes: scratch of isob200_splat_occ_backward_ws_bytes(N,H,W,total_points) bytes enables the sparse
//             forms (per-tile lists of the non-zero gradient pixels; mode 0: points binned by tile, one CTA per
//             tile with the neighbourhood's records staged in shared memory; mode 1: warp per point, sparse /
//             dense per tile); NULL selects the plain window sweep.  Same sums up to fp32 summation order.
int isob200_splat_occ_backward(const float* points, const float* radii, const unsigned char* visible,
                               const int64_t* first_idx, const int64_t* num_points, const float* rs,
                               float radii_s, const float* grad_occ, int N, int H, int W,
                               long long total_points, int mode, float* grad_out, int out_stride,
                               void* ws, size_t ws_bytes, void* stream_) {
  const long long max_points_per_cloud = total_points;
  cudaStream_t st = (cudaStream_t)stream_;
  ISO_CHECK_ARG(N >= 0 && H > 0 && W > 0 && out_stride >= 2, "splat_occ_backward: bad sizes");
  ISO_CHECK_ARG(mode == 0 || mode == 1, "splat_occ_backward: mode must be 0 (fast) or 1 (slow)");
  if (N == 0 || max_points_per_cloud <= 0) return ISOB200_OK;
  ISO_CHECK_ARG(points && radii && first_idx && num_points && grad_occ && grad_out, "splat_occ_backward: null pointer");
  ISO_CHECK_ARG(mode == 1 || rs, "splat_occ_backward: null rs");
  long long need = (max_points_per_cloud + 7) / 8;          // 8 warps (points) per CTA
  int bx = (int)min(need, (long long)kNumSMs * 8 * 8);
  if (N > 1) bx = max(1, min(bx, (kNumSMs * 8 * 8 + N - 1) / N));
  if (ws != nullptr) {
    if (ws_bytes < isob200_splat_occ_backward_ws_bytes(N, H, W, total_points)) {
      set_error("splat_occ_backward: workspace too small");
      return ISOB200_ERR_WORKSPACE;
    }
    const int TX = div_up(W, GT), TY = div_up(H, GT);
    const int nt = N * TX * TY;
    char* base = (char*)ws;
    int* tcnt = (int*)base; base += align_up((size_t)nt * 4);
    int* toff = (int*)base; base += align_up((size_t)nt * 4);
    int* pcnt = (int*)base; base += align_up((size_t)nt * 4);
    int* poff = (int*)base; base += align_up((size_t)nt * 4);
    void* sws = base; const size_t sws_bytes = align_up(scan_ws_bytes(nt, 1)); base += sws_bytes;
    float4* recs = (float4*)base; base += align_up((size_t)N * H * W * sizeof(float4));
    int* plist = (int*)base;
    // small tables: the counting kernels' last CTAs do the scans (two tickets at the head of the scan scratch)
    const bool tail_scan = nt <= LAST_CTA_SCAN_MAX;
    int* tickets = (int*)sws;
    int rc = ISOB200_OK;
    if (tail_scan) ISO_CUDA(cudaMemsetAsync(tickets, 0, 2 * sizeof(int), st));
    gradpix_count_kernel<<<nt, 256, 0, st>>>(grad_occ, H, W, TX, TY, tcnt, tail_scan ? tickets : nullptr, toff);
    ISO_CHECK_LAUNCH("gradpix_count_kernel");
    if (!tail_scan) {
      rc = exclusive_scan_i32(tcnt, toff, nt, 1, nt, nt, sws, sws_bytes, st);
      if (rc) return rc;
    }
    gradpix_fill_kernel<<<nt, 256, 0, st>>>(grad_occ, H, W, TX, TY, toff, recs);
    ISO_CHECK_LAUNCH("gradpix_fill_kernel");
    if (mode == 0) {
      // points binned by the output tile of their centre, then one CTA per (view, tile)
      int pbx = grid_for(max(max_points_per_cloud, 1ll), 256, 8);
      if (N > 1) pbx = max(1, min(pbx, (kNumSMs * 8 + N - 1) / N));
      ISO_CUDA(cudaMemsetAsync(pcnt, 0, (size_t)nt * 4, st));
      if (tail_scan && TX * TY <= OCC_PRIV_MAX_TILES && max_points_per_cloud / N >= OCC_PRIV_CHUNK) {
        const int cbx = (int)min((long long)div_up(max_points_per_cloud / N, OCC_PRIV_CHUNK), (long long)kNumSMs * 4);
        occ_point_count_priv_kernel<<<dim3(cbx, N), 256, (size_t)TX * TY * sizeof(int), st>>>(
            points, visible, first_idx, num_points, H, W, TX, TY, pcnt, tickets + 1, poff, grad_out, out_stride);
        ISO_CHECK_LAUNCH("occ_point_count_priv_kernel");
      } else {
        occ_point_bin_kernel<false><<<dim3(pbx, N), 256, 0, st>>>(points, visible, first_idx, num_points, H, W, TX,
                                                                 TY, pcnt, nullptr, nullptr, grad_out, out_stride);
        ISO_CHECK_LAUNCH("occ_point_bin_kernel<count>");
        rc = exclusive_scan_i32(pcnt, poff, nt, 1, nt, nt, sws, sws_bytes, st);
        if (rc) return rc;
      }
      ISO_CUDA(cudaMemsetAsync(pcnt, 0, (size_t)nt * 4, st));
      occ_point_bin_kernel<true><<<dim3(pbx, N), 256, 0, st>>>(points, visible, first_idx, num_points, H, W, TX, TY,
                                                              pcnt, poff, plist, grad_out, out_stride);
      ISO_CHECK_LAUNCH("occ_point_bin_kernel<fill>");
      splat_occ_backward_tiled_kernel<<<dim3(TX * TY, N), 256, 0, st>>>(points, radii, first_idx, rs, tcnt, toff, recs,
                                                                       pcnt, poff, plist, H, W, TX, TY, grad_out,
                                                                       out_stride);
      ISO_CHECK_LAUNCH("splat_occ_backward_tiled_kernel");
      return ISOB200_OK;
    }
    splat_occ_backward_hybrid_kernel<1><<<dim3(bx, N), 256, 0, st>>>(points, radii, visible, first_idx, num_points,
                                                                      rs, radii_s, grad_occ, tcnt, toff, recs, H, W,
                                                                      TX, TY, grad_out, out_stride);
    ISO_CHECK_LAUNCH("splat_occ_backward_hybrid_kernel");
    return ISOB200_OK;
  }
  if (mode == 0)
    splat_occ_backward_kernel<0><<<dim3(bx, N), 256, 0, st>>>(points, radii, visible, first_idx, num_points, rs,
                                                             radii_s, grad_occ, H, W, grad_out, out_stride);
  else
    splat_occ_backward_kernel<1><<<dim3(bx, N), 256, 0, st>>>(points, radii, visible, first_idx, num_points, rs,
                                                             radii_s, grad_occ, H, W, grad_out, out_stride);
  ISO_CHECK_LAUNCH("splat_occ_backward_kernel");
  return ISOB200_OK;
}

size_t isob200_splat_search_radius_ws_bytes(int N) { return align_up((size_t)N * (256 + 4) * sizeof(unsigned)); }

// rs[n] = median(radii[visible rows of view n].flatten()) * radii_s  (rasterizer.py:881-884), 0 for a view
// without visible points.  visible may be NULL (all rows).
int isob200_splat_search_radius(const float* radii, const unsigned char* visible, const int64_t* first_idx,
                                const int64_t* num_points, int N, long long max_points_per_cloud, float radii_s,
                                float* rs, void* ws, size_t ws_bytes, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  if (N <= 0) return ISOB200_OK;
  ISO_CHECK_ARG(radii && first_idx && num_points && rs && ws, "splat_search_radius: null pointer");
  if (ws_bytes < isob200_splat_search_radius_ws_bytes(N)) {
    set_error("splat_search_radius: workspace too small");
    return ISOB200_ERR_WORKSPACE;
  }
  unsigned* hist = (unsigned*)ws;
  unsigned* state = hist + (size_t)N * 256;
  ISO_CUDA(cudaMemsetAsync(ws, 0, (size_t)N * (256 + 4) * sizeof(unsigned), st));
  int bx = grid_for(div_up(max(max_points_per_cloud, 1ll), 4), 256, 8);   // four points per thread and trip
  if (N > 1) bx = max(1, min(bx, (kNumSMs * 8 + N - 1) / N));
  for (int pass = 0; pass < 4; ++pass) {
    const int shift = 24 - 8 * pass;
    median_pass_kernel<<<dim3(bx, N), 256, 0, st>>>(radii, visible, first_idx, num_points, shift, pass == 0,
                                                    pass == 3, radii_s, state, hist, rs);
    ISO_CHECK_LAUNCH("median_pass_kernel");
  }
  return ISOB200_OK;
}

// == DSS._C._backward_zbuf (rasterize_points.h:388-419): in place, z_grad[p*stride] += ...
int isob200_splat_zbuf_backward(const int* idx, const float* grad_zbuf, int N, int H, int W, int K,
                                float* z_grad, int stride, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  const long long npix = (long long)N * H * W;
  if (npix == 0 || K == 0) return ISOB200_OK;
  ISO_CHECK_ARG(idx && grad_zbuf && z_grad && stride >= 1, "splat_zbuf_backward: bad argument");
  splat_zbuf_backward_kernel<<<grid_for(npix, 256, 8), 256, 0, st>>>(idx, grad_zbuf, npix, K, z_grad, stride);
  ISO_CHECK_LAUNCH("splat_zbuf_backward_kernel");
  return ISOB200_OK;
}

// visible (P,) uint8 must be zero-initialised by the caller.
int isob200_splat_visibility(const int* idx, const float* mask, long long npix, int K, long long P,
                             unsigned char* visible, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  if (npix == 0 || K == 0 || P == 0) return ISOB200_OK;
  ISO_CHECK_ARG(idx && visible, "splat_visibility: null pointer");
  splat_visibility_kernel<<<grid_for(npix * K, 256, 8), 256, 0, st>>>(idx, mask, npix, K, P, visible);
  ISO_CHECK_LAUNCH("splat_visibility_kernel");
  return ISOB200_OK;
}

// out (npix, C+1) = [normalised weighted colour, occ]; C in 1..4.  weights_out may be NULL.
int isob200_splat_blend(const int* idx, const float* qvalue, const float* occ, const float* scaler,
                        const float* feat, int feat_stride, long long npix, int K, int C, float eps,
                        float* out, float* weights_out, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  if (npix == 0) return ISOB200_OK;
  ISO_CHECK_ARG(idx && qvalue && occ && feat && out, "splat_blend: null pointer");
  ISO_CHECK_ARG(C >= 1 && C <= 4 && feat_stride >= C, "splat_blend: C must be in [1,4]");
  const int g = grid_for(npix, 256, 8);
#define BL(CC) case CC: splat_blend_kernel<CC><<<g, 256, 0, st>>>(idx, qvalue, occ, scaler, feat, feat_stride, npix, K, eps, out, weights_out); break;
  switch (C) { BL(1) BL(2) BL(3) BL(4) }
#undef BL
  ISO_CHECK_LAUNCH("splat_blend_kernel");
  return ISOB200_OK;
}

// grad_feat (P, feat_stride) += ; must be zero-initialised by the caller.
int isob200_splat_blend_backward(const int* idx, const float* weights, const float* grad_out,
                                 long long npix, int K, int C, float eps, float* grad_feat,
                                 int feat_stride, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  if (npix == 0) return ISOB200_OK;
  ISO_CHECK_ARG(idx && weights && grad_out && grad_feat, "splat_blend_backward: null pointer");
  ISO_CHECK_ARG(C >= 1 && C <= 4 && feat_stride >= C, "splat_blend_backward: C must be in [1,4]");
  const int g = grid_for(npix, 256, 8);
#define BB(CC) case CC: splat_blend_backward_kernel<CC><<<g, 256, 0, st>>>(idx, weights, grad_out, npix, K, eps, grad_feat, feat_stride); break;
  switch (C) { BB(1) BB(2) BB(3) BB(4) }
#undef BB
  ISO_CHECK_LAUNCH("splat_blend_backward_kernel");
  return ISOB200_OK;
}

}  // extern "C"
