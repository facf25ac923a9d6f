// isob200 -- point-to-ray search for the in-surface sampler.
//
// Model.sample_offsurface_using_isopoints (DSS/models/combined_modeling.py:325-352) bounds the sampling
// segment of every camera ray by the visible iso-point closest to the ray, once for the front-facing and
// once for the occluded iso-points.  The reference materialises two (R, M) matrices per view
//     ray_sq      = ((p - C) . d)^2
//     dist_to_ray = |p - C|^2 - ray_sq
// and takes topk(k = 1, smallest) of dist_to_ray (marked "TODO: faster search").  Here one warp owns one
// ray, a CTA of 8 warps stages tiles of points in shared memory (each tile is read from L2 once per 8
// rays), every lane keeps the running (dist, index) minimum of its stride and a shuffle reduction picks
// the winner; ties go to the lower index.  Nothing of size R x M ever exists.
//
// Arithmetic: separate fp32 multiplies and adds in the reference's order (no FMA contraction), so that the
// selected point agrees with the dense computation except on exact ties.
#include "common.cuh"
#include "isob200.h"

namespace isob200 {

constexpr int RAY_WARPS = 8;          // rays per CTA
constexpr int RAY_TILE = 2048;        // points per shared-memory tile (24 KB)

__global__ void __launch_bounds__(RAY_WARPS * 32)
ray_nearest_point_kernel(const float* __restrict__ origins, int origin_stride, const float* __restrict__ dirs,
                         int R, const float* __restrict__ points, int M, float* __restrict__ t_sq_out,
                         float* __restrict__ dist_out, int* __restrict__ idx_out) {
  __shared__ float sx[RAY_TILE], sy[RAY_TILE], sz[RAY_TILE];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ray = blockIdx.x * RAY_WARPS + warp;
  const bool live = ray < R;
  float ox = 0.f, oy = 0.f, oz = 0.f, dx = 0.f, dy = 0.f, dz = 0.f;
  if (live) {
    const float* o = origins + (size_t)ray * origin_stride;   // stride 0: one origin for all rays
    ox = o[0]; oy = o[1]; oz = o[2];
    dx = dirs[(size_t)ray * 3 + 0]; dy = dirs[(size_t)ray * 3 + 1]; dz = dirs[(size_t)ray * 3 + 2];
  }
  float best = __int_as_float(0x7f800000);   // +inf
  float best_t = 0.f;
  int best_i = 0x7fffffff;
  for (int base = 0; base < M; base += RAY_TILE) {
    const int n = min(RAY_TILE, M - base);
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += RAY_WARPS * 32) {
      const float* p = points + (size_t)(base + i) * 3;
      sx[i] = p[0]; sy[i] = p[1]; sz[i] = p[2];
    }
    __syncthreads();
    if (live) {
      for (int i = lane; i < n; i += 32) {
        const float px = __fsub_rn(sx[i], ox), py = __fsub_rn(sy[i], oy), pz = __fsub_rn(sz[i], oz);
        const float t = __fadd_rn(__fadd_rn(__fmul_rn(px, dx), __fmul_rn(py, dy)), __fmul_rn(pz, dz));
        const float tt = __fmul_rn(t, t);
        const float pp = __fadd_rn(__fadd_rn(__fmul_rn(px, px), __fmul_rn(py, py)), __fmul_rn(pz, pz));
        const float d = __fsub_rn(pp, tt);
        // strict <: within a lane indices only grow, so the first minimum is kept; NaN never wins
        if (d < best) { best = d; best_t = tt; best_i = base + i; }
      }
    }
  }
  if (!live) return;
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    const float od = __shfl_xor_sync(0xffffffffu, best, off);
    const float ot = __shfl_xor_sync(0xffffffffu, best_t, off);
    const int oi = __shfl_xor_sync(0xffffffffu, best_i, off);
    if (od < best || (od == best && oi < best_i)) { best = od; best_t = ot; best_i = oi; }
  }
  if (lane == 0) {
    const bool found = best_i != 0x7fffffff;
    if (t_sq_out) t_sq_out[ray] = found ? best_t : 0.f;
    if (dist_out) dist_out[ray] = best;
    if (idx_out) idx_out[ray] = found ? best_i : -1;
  }
}

}  // namespace isob200

using namespace isob200;

extern "C" int isob200_ray_nearest_point(const float* origins, int n_origins, const float* dirs, int R,
                                         const float* points, int M, float* t_sq, float* dist, int* idx,
                                         void* stream) {
  ISO_CHECK_ARG(R >= 0 && M >= 0, "ray_nearest_point: negative size");
  if (R == 0) return ISOB200_OK;
  ISO_CHECK_ARG(origins && dirs, "ray_nearest_point: null rays");
  ISO_CHECK_ARG(n_origins == 1 || n_origins == R, "ray_nearest_point: %d origins for %d rays (1 or R)", n_origins, R);
  ISO_CHECK_ARG(M == 0 || points, "ray_nearest_point: null points");
  ISO_CHECK_ARG(t_sq || dist || idx, "ray_nearest_point: no output");
  const int grid = div_up(R, RAY_WARPS);
  ray_nearest_point_kernel<<<grid, RAY_WARPS * 32, 0, (cudaStream_t)stream>>>(
      origins, n_origins == 1 ? 0 : 3, dirs, R, points, M, t_sq, dist, idx);
  ISO_CHECK_LAUNCH("ray_nearest_point_kernel");
  return ISOB200_OK;
}
