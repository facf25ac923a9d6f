// FRNN fixed-radius K-nearest query, gather and backward for sm_100a.
//
// Replaces external/FRNN/frnn/csrc/grid/grid.cu:186-440 (FindNbrs{2,3}DKernel<K>: one thread
// per query, K-entry MinK in local memory, fixed <<<256,256>>> grid), frnn.py:304-352
// (frnn_gather as an expand+gather+mask chain) and backward/backward.cu:8-147.
//
// Three query kernels with identical results.  Dense grids (>= 1 point per cell) and K <= 20: one THREAD per
// query, collect-then-select (frnn_query_collect_kernel below: trial radius from the local density, candidates
// appended to a per-thread shared-memory column, K smallest emitted; warp-cooperative exact fallback).  Otherwise
// the group kernels: a GROUP of GW lanes (GW = 8/16/32 >= K) cooperates on one query.
//  * the candidate cells of a query form (2c+1)^(D-1) runs that are CONTIGUOUS in the sorted
//    point array (z is the fastest cell index), so the group streams each run with coalesced
//    loads, GW candidates per step -- no per-thread divergent pointer chasing;
//  * the running K-best list lives one entry per lane, kept sorted by (dist, index); an
//    insertion is a ballot + shuffle-up, no local memory, no bubble sort at the end;
//  * results are written straight to the caller's (N,P1,K) tensors (int64 or int32 indices),
//    including the -1 padding, so no fill pass is needed.
// Distances use the reference's exact fp32 expression as nvcc contracts it (SASS of grid.cu:
// FMUL dy*dy; FFMA dx; FFMA dz), spelled with intrinsics so it can never be re-contracted.
// Equal distances are ordered by smaller original index (the reference keeps whichever it saw
// first, which depends on its atomic insertion order -- mink.cuh:64).
#include "common.cuh"
#include <float.h>
#include <limits.h>

namespace isob200 {

template <int D>
__device__ __forceinline__ float sqdist_ref(const float* __restrict__ p2, const float* q) {
  // SASS of FindNbrs{2,3}DKernel (every K): FMUL dy*dy ; FFMA dx*dx + . ; FFMA dz*dz + .
  const float dx = __fsub_rn(p2[0], q[0]);
  const float dy = __fsub_rn(p2[1], q[1]);
  float s = __fmaf_rn(dx, dx, __fmul_rn(dy, dy));
  if (D == 3) {
    const float dz = __fsub_rn(p2[2], q[2]);
    s = __fmaf_rn(dz, dz, s);
  }
  return s;
}

// Squared distance (in world units, conservative: never larger than the true minimum) from the
// query at cell-space coordinate qc to cell index i along one axis.  The slack of 1e-3 cell
// absorbs the fp32 rounding of (p - min) * delta that assigned points to cells.
__device__ __forceinline__ float axis_gap2(float qc, int i, float inv_delta) {
  const float g = fmaxf(fmaxf((float)i - qc, qc - (float)(i + 1)), 0.0f);
  const float gs = fmaxf(g - 1e-3f, 0.0f) * inv_delta;
  return gs * gs;
}

template <int D, int GW, typename IdxT>
__global__ void __launch_bounds__(256)
frnn_query_kernel(const float* __restrict__ q_points,     // (N,P1,D) in processing order
                  const int* __restrict__ q_order,        // (N,P1) processing slot -> original row, or null
                  const int64_t* __restrict__ lengths1,   // (N,) or null
                  const int64_t* __restrict__ lengths2,   // (N,) or null
                  const float* __restrict__ sorted_points2,  // (N,P2,D)
                  const int* __restrict__ cell_off2,         // (N,G)
                  const int* __restrict__ sorted_idxs2,      // (N,P2)
                  const float* __restrict__ params, const float* __restrict__ rs, int N, int P1,
                  int P2, int G, int K, float* __restrict__ dists, IdxT* __restrict__ idxs) {
  constexpr int PS = (D == 3) ? ISO_G3_SIZE : ISO_G2_SIZE;
  constexpr int GROUPS = 256 / GW;
  const int lane = threadIdx.x & 31;
  const int gl = lane & (GW - 1);                       // lane within group
  const unsigned gshift = lane & ~(GW - 1);             // first warp lane of this group
  const unsigned gmask = (GW == 32) ? 0xffffffffu : (((1u << GW) - 1u) << gshift);
  const long long total = (long long)N * P1;
  const long long ngroups = (long long)gridDim.x * GROUPS;

  for (long long item = (long long)blockIdx.x * GROUPS + threadIdx.x / GW; item < total;
       item += ngroups) {
    const int n = (int)(item / P1);
    const int s = (int)(item - (long long)n * P1);
    const int len1 = lengths1 ? (int)min((long long)lengths1[n], (long long)P1) : P1;
    if (s >= len1) {  // padded row: reference leaves the -1 fill (grid.cu:422-423)
      if (gl < K) {
        const size_t o = ((size_t)n * P1 + s) * K + gl;
        dists[o] = -1.f;
        idxs[o] = (IdxT)-1;
      }
      continue;
    }
    const int row = q_order ? q_order[(size_t)n * P1 + s] : s;
    const int len2 = lengths2 ? (int)min((long long)lengths2[n], (long long)P2) : P2;
    const float* prm = params + (size_t)n * PS;
    const float r = rs[n];
    const float r2 = __fmul_rn(r, r);
    const float delta = prm[D];
    const float inv_delta = 1.0f / delta;
    float q[D], qc[D];
#pragma unroll
    for (int d = 0; d < D; ++d) q[d] = q_points[((size_t)n * P1 + s) * D + d];

    // candidate cell range, grid.cu:305-316 (fp32: (p - min -/+ r) * delta, floor)
    int lo[D], hi[D], res[D], cq[D];
    bool nonempty = true;
#pragma unroll
    for (int d = 0; d < D; ++d) {
      const float rel = __fsub_rn(q[d], prm[d]);
      res[d] = (int)prm[D + 1 + d];
      lo[d] = max(__float2int_rd(__fmul_rn(__fsub_rn(rel, r), delta)), 0);
      hi[d] = min(__float2int_rd(__fmul_rn(__fadd_rn(rel, r), delta)), res[d] - 1);
      nonempty = nonempty && (lo[d] <= hi[d]);
      qc[d] = rel * delta;
      cq[d] = min(max(__float2int_rd(qc[d]), lo[d]), hi[d]);
    }
    const int grid_total = (int)prm[2 * D + 1];

    float best_d = FLT_MAX;   // lane gl holds the gl-th best (dist, idx), ascending
    int best_i = INT_MAX;
    float bound = r2;         // min(r2, current K-th best distance): group-uniform pruning bound

    const float* pts2 = sorted_points2 + (size_t)n * P2 * D;
    const int* off2 = cell_off2 + (size_t)n * G;
    const int* sid2 = sorted_idxs2 + (size_t)n * P2;

    if (nonempty) {
      // Best-first traversal: rings of runs around the query's own cell.  Once K candidates are
      // held, a run (and the far cells of a run) whose closest possible point is beyond the K-th
      // best distance cannot change the result (strict >, so equal-distance ties still compete)
      // and is skipped.  The visited set shrinks from (2r/cell+1)^D cells to those overlapping the
      // K-th-neighbour ball; the result is the same K smallest (dist, index) pairs.
      const int R0 = max(cq[0] - lo[0], hi[0] - cq[0]);
      const int R1 = (D == 3) ? max(cq[1] - lo[1], hi[1] - cq[1]) : 0;
      const int R = max(R0, R1);
      for (int ring = 0; ring <= R; ++ring) {
        const int xstep = (D == 2) ? max(2 * ring, 1) : 1;   // 2-D: the ring is just the two end cells
        for (int ox = -ring; ox <= ring; ox += xstep) {
          const int x = cq[0] + ox;
          if (x < lo[0] || x > hi[0]) continue;
          const float gx2 = axis_gap2(qc[0], x, inv_delta);
          if (gx2 > bound) continue;
          const int ystep = (D == 3 && (ox == -ring || ox == ring)) ? 1 : max(2 * ring, 1);
          for (int oy = (D == 3 ? -ring : 0); oy <= (D == 3 ? ring : 0); oy += ystep) {
            int c0;
            float gxy2 = gx2;
            int zlo, zhi;       // cell range along the fastest axis
            float qcz;
            if (D == 3) {
              const int y = cq[1] + oy;
              if (y < lo[1] || y > hi[1]) continue;
              gxy2 += axis_gap2(qc[1], y, inv_delta);
              if (gxy2 > bound) continue;
              zlo = lo[2]; zhi = hi[2]; qcz = qc[2];
              c0 = (x * res[1] + y) * res[2];
            } else {
              zlo = lo[1]; zhi = hi[1]; qcz = qc[1];
              c0 = x * res[1];
            }
            // scan the cells [za, zb] of this run (contiguous in the sorted array)
            auto scan = [&](int za, int zb) {
              if (za > zb) return;
              const int start = off2[c0 + za];
              const int end = (c0 + zb + 1 == grid_total) ? len2 : off2[c0 + zb + 1];
              for (int base = start; base < end; base += GW) {
                const int j = base + gl;
                const bool valid = j < end;
                float d = FLT_MAX;
                if (valid) d = sqdist_ref<D>(pts2 + (size_t)j * D, q);
                const bool cand = valid && (d <= bound);
                unsigned m = __ballot_sync(gmask, cand) & gmask;
                if (m == 0) continue;
                int ci_mine = cand ? sid2[j] : INT_MAX;
                while (m) {
                  const int src = __ffs(m) - 1;
                  m &= m - 1;
                  const float cd = __shfl_sync(gmask, d, src);
                  const int ci = __shfl_sync(gmask, ci_mine, src);
                  const bool less = (best_d < cd) || (best_d == cd && best_i < ci);
                  const int pos = __popc(__ballot_sync(gmask, less) & gmask);
                  const float up_d = __shfl_up_sync(gmask, best_d, 1, GW);
                  const int up_i = __shfl_up_sync(gmask, best_i, 1, GW);
                  if (pos < K) {
                    if (gl > pos) { best_d = up_d; best_i = up_i; }
                    else if (gl == pos) { best_d = cd; best_i = ci; }
                  }
                }
                bound = fminf(r2, __shfl_sync(gmask, best_d, (int)gshift + K - 1));
              }
            };
            // cells of the run that can still hold a candidate within `bound`
            auto zrange = [&](int& a, int& b) {
              const float gz = sqrtf(fmaxf(bound - gxy2, 0.0f)) * delta + 2e-3f;
              a = max(zlo, __float2int_ru(qcz - 1.0f - gz));
              b = min(zhi, __float2int_rd(qcz + gz));
            };
            int za, zb;
            zrange(za, zb);
            if (ring == 0) {
              // the query's own run: its three central cells first, so that the K-th distance bound
              // exists before the far cells (and every other run) are looked at
              const int zc = min(max(__float2int_rd(qcz), za), zb);
              const int wa = max(za, zc - 1), wb = min(zb, zc + 1);
              scan(wa, wb);
              zrange(za, zb);
              scan(za, min(zb, wa - 1));
              scan(max(za, wb + 1), zb);
            } else {
              scan(za, zb);
            }
          }
        }
      }
    }
    if (gl < K) {
      const size_t o = ((size_t)n * P1 + row) * K + gl;
      const bool found = best_i != INT_MAX;
      dists[o] = found ? best_d : -1.f;
      idxs[o] = found ? (IdxT)best_i : (IdxT)-1;
    }
  }
}

// Block traversal of the (2c+1)^D cell block: the lanes prefetch every run's slice and its distance
// to the query, the runs next to the query are scanned first and the others are skipped when they
// cannot beat the K-th best.  The better kernel when the grid is mostly EMPTY cells (surface-like clouds with a small cell: at
// BASELINE config 2 there are ~0.1 points per cell), where skipping an empty run costs two offset loads
// and the pruning bookkeeping of the kernel above costs more than it saves.
template <int D, int GW, typename IdxT>
__global__ void __launch_bounds__(256)
frnn_query_exhaustive_kernel(const float* __restrict__ q_points,     // (N,P1,D) in processing order
                  const int* __restrict__ q_order,        // (N,P1) processing slot -> original row, or null
                  const int64_t* __restrict__ lengths1,   // (N,) or null
                  const int64_t* __restrict__ lengths2,   // (N,) or null
                  const float* __restrict__ sorted_points2,  // (N,P2,D)
                  const int* __restrict__ cell_off2,         // (N,G)
                  const int* __restrict__ sorted_idxs2,      // (N,P2)
                  const float* __restrict__ params, const float* __restrict__ rs, int N, int P1,
                  int P2, int G, int K, float* __restrict__ dists, IdxT* __restrict__ idxs) {
  constexpr int PS = (D == 3) ? ISO_G3_SIZE : ISO_G2_SIZE;
  constexpr int GROUPS = 256 / GW;
  const int lane = threadIdx.x & 31;
  const int gl = lane & (GW - 1);                       // lane within group
  const unsigned gshift = lane & ~(GW - 1);             // first warp lane of this group
  const unsigned gmask = (GW == 32) ? 0xffffffffu : (((1u << GW) - 1u) << gshift);
  const long long total = (long long)N * P1;
  const long long ngroups = (long long)gridDim.x * GROUPS;

  for (long long item = (long long)blockIdx.x * GROUPS + threadIdx.x / GW; item < total;
       item += ngroups) {
    const int n = (int)(item / P1);
    const int s = (int)(item - (long long)n * P1);
    const int len1 = lengths1 ? (int)min((long long)lengths1[n], (long long)P1) : P1;
    if (s >= len1) {  // padded row: reference leaves the -1 fill (grid.cu:422-423)
      if (gl < K) {
        const size_t o = ((size_t)n * P1 + s) * K + gl;
        dists[o] = -1.f;
        idxs[o] = (IdxT)-1;
      }
      continue;
    }
    const int row = q_order ? q_order[(size_t)n * P1 + s] : s;
    const int len2 = lengths2 ? (int)min((long long)lengths2[n], (long long)P2) : P2;
    const float* prm = params + (size_t)n * PS;
    const float r = rs[n];
    const float r2 = __fmul_rn(r, r);
    float q[D];
#pragma unroll
    for (int d = 0; d < D; ++d) q[d] = q_points[((size_t)n * P1 + s) * D + d];

    // candidate cell range, grid.cu:305-316 (fp32: (p - min -/+ r) * delta, floor)
    int lo[D], hi[D], res[D];
    float qc[D];
    const float inv_delta = 1.0f / prm[D];
#pragma unroll
    for (int d = 0; d < D; ++d) {
      const float rel = __fsub_rn(q[d], prm[d]);
      const float delta = prm[D];
      res[d] = (int)prm[D + 1 + d];
      lo[d] = max(__float2int_rd(__fmul_rn(__fsub_rn(rel, r), delta)), 0);
      hi[d] = min(__float2int_rd(__fmul_rn(__fadd_rn(rel, r), delta)), res[d] - 1);
      qc[d] = rel * delta;
    }
    const int grid_total = (int)prm[2 * D + 1];

    float best_d = FLT_MAX;   // lane gl holds the gl-th best (dist, idx), ascending
    int best_i = INT_MAX;
    float kth_d = FLT_MAX;    // current K-th best distance (pruning bound), group-uniform

    const float* pts2 = sorted_points2 + (size_t)n * P2 * D;
    const int* off2 = cell_off2 + (size_t)n * G;
    const int* sid2 = sorted_idxs2 + (size_t)n * P2;

    bool nonempty = true;
#pragma unroll
    for (int d = 0; d < D; ++d) nonempty = nonempty && (lo[d] <= hi[d]);
    if (nonempty) {
      const int ny = (D == 3) ? (hi[1] - lo[1] + 1) : 1;
      const int nruns = (hi[0] - lo[0] + 1) * ny;  // <= 0 when empty
      // the [start, end) slices of up to GW runs are fetched by the lanes in parallel (one division
      // and two offset loads per lane), then handed round with shuffles: an empty run -- most of
      // them on a surface -- costs two shuffles and a compare instead of a dependent load chain
      for (int rbase = 0; rbase < nruns; rbase += GW) {
       int my_start = 0, my_end = 0;
       float my_gap2 = 0.f;   // conservative squared distance from the query to the run's cell column
       {
        const int run = rbase + gl;
        if (run < nruns) {
          int c0, c1;
          if (D == 3) {
            const int x = lo[0] + run / ny, y = lo[1] + run % ny;
            c0 = (x * res[1] + y) * res[2] + lo[2];
            c1 = (x * res[1] + y) * res[2] + hi[2];
            my_gap2 = axis_gap2(qc[0], x, inv_delta) + axis_gap2(qc[1], y, inv_delta);
          } else {
            const int x = lo[0] + run;
            c0 = x * res[1] + lo[1];
            c1 = x * res[1] + hi[1];
            my_gap2 = axis_gap2(qc[0], x, inv_delta);
          }
          my_start = off2[c0];
          my_end = (c1 + 1 == grid_total) ? len2 : off2[c1 + 1];
        }
       }
       const int nr = min(GW, nruns - rbase);
       // two sweeps over the prefetched runs: the ones next to the query first (they establish the
       // K-th distance), then the rest -- skipped when even their nearest cell wall is farther than the
       // K-th best (strict >, equal-distance ties still compete)
       const float near2 = inv_delta * inv_delta;
       for (int rj2 = 0; rj2 < 2 * nr; ++rj2) {
        const int rj = rj2 < nr ? rj2 : rj2 - nr;
        const float gap2 = __shfl_sync(gmask, my_gap2, (int)gshift + rj);
        if ((gap2 <= near2) != (rj2 < nr)) continue;
        if (gap2 > fminf(r2, kth_d)) continue;
        const int start = __shfl_sync(gmask, my_start, (int)gshift + rj);
        const int end = __shfl_sync(gmask, my_end, (int)gshift + rj);
        for (int base = start; base < end; base += GW) {
          const int j = base + gl;
          const bool valid = j < end;
          float d = FLT_MAX;
          if (valid) d = sqdist_ref<D>(pts2 + (size_t)j * D, q);
          const bool cand = valid && (d <= r2) && (d <= kth_d);
          unsigned m = __ballot_sync(gmask, cand) & gmask;
          if (m == 0) continue;
          int ci_mine = cand ? sid2[j] : INT_MAX;
          while (m) {
            const int src = __ffs(m) - 1;
            m &= m - 1;
            const float cd = __shfl_sync(gmask, d, src);
            const int ci = __shfl_sync(gmask, ci_mine, src);
            const bool less = (best_d < cd) || (best_d == cd && best_i < ci);
            const int pos = __popc(__ballot_sync(gmask, less) & gmask);
            const float up_d = __shfl_up_sync(gmask, best_d, 1, GW);
            const int up_i = __shfl_up_sync(gmask, best_i, 1, GW);
            if (pos < K) {
              if (gl > pos) { best_d = up_d; best_i = up_i; }
              else if (gl == pos) { best_d = cd; best_i = ci; }
            }
          }
          kth_d = __shfl_sync(gmask, best_d, (int)gshift + K - 1);
        }
       }
      }
    }
    if (gl < K) {
      const size_t o = ((size_t)n * P1 + row) * K + gl;
      const bool found = best_i != INT_MAX;
      dists[o] = found ? best_d : -1.f;
      idxs[o] = found ? (IdxT)best_i : (IdxT)-1;
    }
  }
}

// ---- dense grids: one THREAD per query, collect-then-select ---------------------------------------------
// The group-cooperative kernels above spend ~2 300 warp instructions per pair of queries at BASELINE config 3
// (500 k uniform points, K = 16, ~8 points per cell): a third on the ring / run bookkeeping, a third on 16-wide
// candidate steps that are mostly predicated off, a third on serial ballot + shuffle insertions -- and the two
// groups of a warp diverge.  When the grid is dense (>= 1 point per cell) the K-th neighbour lies well inside the
// search radius, so here each thread FIRST fixes a trial radius from the local density (the candidate block's
// point count is two offset loads per run) such that the ball holds lambda = K + 4 sqrt(K) points on average,
// THEN makes one pass over the cells that ball touches and appends every candidate inside it to a per-thread
// column of shared memory (no ordering work at all: a predicated 8-byte store), and finally reads its <= 48
// entries back K times to emit the K smallest (dist, index) keys in order.  Exactness does not rest on the
// estimate: a thread whose ball held fewer than K (or more than the column's capacity of) candidates hands its
// query to the warp, which runs the pruned best-first search of frnn_query_kernel on it with all 32 lanes.
// Results are identical to the other kernels (same fp32 distance expression, same (dist, index) order).
constexpr int COLLECT_THREADS = 128;
constexpr int COLLECT_CAP = 48;       // entries per thread: 48 x 128 x 8 B = 48 KB of shared memory per CTA

// The pruned best-first search of frnn_query_kernel for ONE query, run by a full warp (lane l ends up holding
// the l-th best (dist, index), ascending; FLT_MAX / INT_MAX where fewer were found).
template <int D>
__device__ __forceinline__ void warp_pruned_search(const float (&q)[D], const float* __restrict__ prm, float r2,
                                                   float r, const float* __restrict__ pts2,
                                                   const int* __restrict__ off2, const int* __restrict__ sid2,
                                                   int len2, int K, float& best_d, int& best_i) {
  const int gl = threadIdx.x & 31;
  const float delta = prm[D];
  const float inv_delta = 1.0f / delta;
  int lo[D], hi[D], res[D], cq[D];
  float qc[D];
  bool nonempty = true;
#pragma unroll
  for (int d = 0; d < D; ++d) {
    const float rel = __fsub_rn(q[d], prm[d]);
    res[d] = (int)prm[D + 1 + d];
    lo[d] = max(__float2int_rd(__fmul_rn(__fsub_rn(rel, r), delta)), 0);
    hi[d] = min(__float2int_rd(__fmul_rn(__fadd_rn(rel, r), delta)), res[d] - 1);
    nonempty = nonempty && (lo[d] <= hi[d]);
    qc[d] = rel * delta;
    cq[d] = min(max(__float2int_rd(qc[d]), lo[d]), hi[d]);
  }
  const int grid_total = (int)prm[2 * D + 1];
  best_d = FLT_MAX;
  best_i = INT_MAX;
  float bound = r2;
  if (!nonempty) return;
  const int R0 = max(cq[0] - lo[0], hi[0] - cq[0]);
  const int R1 = (D == 3) ? max(cq[1] - lo[1], hi[1] - cq[1]) : 0;
  const int R = max(R0, R1);
  for (int ring = 0; ring <= R; ++ring) {
    const int xstep = (D == 2) ? max(2 * ring, 1) : 1;
    for (int ox = -ring; ox <= ring; ox += xstep) {
      const int x = cq[0] + ox;
      if (x < lo[0] || x > hi[0]) continue;
      const float gx2 = axis_gap2(qc[0], x, inv_delta);
      if (gx2 > bound) continue;
      const int ystep = (D == 3 && (ox == -ring || ox == ring)) ? 1 : max(2 * ring, 1);
      for (int oy = (D == 3 ? -ring : 0); oy <= (D == 3 ? ring : 0); oy += ystep) {
        int c0;
        float gxy2 = gx2;
        int zlo, zhi;
        float qcz;
        if (D == 3) {
          const int y = cq[1] + oy;
          if (y < lo[1] || y > hi[1]) continue;
          gxy2 += axis_gap2(qc[1], y, inv_delta);
          if (gxy2 > bound) continue;
          zlo = lo[2]; zhi = hi[2]; qcz = qc[2];
          c0 = (x * res[1] + y) * res[2];
        } else {
          zlo = lo[1]; zhi = hi[1]; qcz = qc[1];
          c0 = x * res[1];
        }
        const float gz = sqrtf(fmaxf(bound - gxy2, 0.0f)) * delta + 2e-3f;
        const int za = max(zlo, __float2int_ru(qcz - 1.0f - gz));
        const int zb = min(zhi, __float2int_rd(qcz + gz));
        if (za > zb) continue;
        const int start = off2[c0 + za];
        const int end = (c0 + zb + 1 == grid_total) ? len2 : off2[c0 + zb + 1];
        for (int base = start; base < end; base += 32) {
          const int j = base + gl;
          const bool valid = j < end;
          float d = FLT_MAX;
          if (valid) d = sqdist_ref<D>(pts2 + (size_t)j * D, q);
          const bool cand = valid && (d <= bound);
          unsigned m = __ballot_sync(0xffffffffu, cand);
          if (m == 0) continue;
          const int ci_mine = cand ? sid2[j] : INT_MAX;
          while (m) {
            const int src = __ffs(m) - 1;
            m &= m - 1;
            const float cd = __shfl_sync(0xffffffffu, d, src);
            const int ci = __shfl_sync(0xffffffffu, ci_mine, src);
            const bool less = (best_d < cd) || (best_d == cd && best_i < ci);
            const int pos = __popc(__ballot_sync(0xffffffffu, less));
            const float up_d = __shfl_up_sync(0xffffffffu, best_d, 1);
            const int up_i = __shfl_up_sync(0xffffffffu, best_i, 1);
            if (pos < K) {
              if (gl > pos) { best_d = up_d; best_i = up_i; }
              else if (gl == pos) { best_d = cd; best_i = ci; }
            }
          }
          bound = fminf(r2, __shfl_sync(0xffffffffu, best_d, K - 1));
        }
      }
    }
  }
}

template <int D, typename IdxT>
__global__ void __launch_bounds__(COLLECT_THREADS)
frnn_query_collect_kernel(const float* __restrict__ q_points, const int* __restrict__ q_order,
                          const int64_t* __restrict__ lengths1, const int64_t* __restrict__ lengths2,
                          const float* __restrict__ sorted_points2, const int* __restrict__ cell_off2,
                          const int* __restrict__ sorted_idxs2, const float* __restrict__ params,
                          const float* __restrict__ rs, int N, int P1, int P2, int G, int K, float lambda,
                          float* __restrict__ dists, IdxT* __restrict__ idxs) {
  constexpr int PS = (D == 3) ? ISO_G3_SIZE : ISO_G2_SIZE;
  extern __shared__ unsigned long long s_keys[];     // [COLLECT_CAP][COLLECT_THREADS]
  const int tid = threadIdx.x, lane = tid & 31;
  const long long total = (long long)N * P1;
  const long long nslots = ((total + COLLECT_THREADS - 1) / COLLECT_THREADS) * COLLECT_THREADS;
  for (long long item = (long long)blockIdx.x * COLLECT_THREADS + tid; item < nslots;
       item += (long long)gridDim.x * COLLECT_THREADS) {
    const bool in_range = item < total;
    const int n = in_range ? (int)(item / P1) : 0;
    const int s = in_range ? (int)(item - (long long)n * P1) : 0;
    const int len1 = lengths1 ? (int)min((long long)lengths1[n], (long long)P1) : P1;
    const bool live = in_range && s < len1;
    const int len2 = lengths2 ? (int)min((long long)lengths2[n], (long long)P2) : P2;
    const float* prm = params + (size_t)n * PS;
    const float r = rs[n];
    const float r2 = __fmul_rn(r, r);
    const float* pts2 = sorted_points2 + (size_t)n * P2 * D;
    const int* off2 = cell_off2 + (size_t)n * G;
    const int* sid2 = sorted_idxs2 + (size_t)n * P2;
    float q[D];
#pragma unroll
    for (int d = 0; d < D; ++d) q[d] = live ? q_points[((size_t)n * P1 + s) * D + d] : 0.f;
    const int row = (live && q_order) ? q_order[(size_t)n * P1 + s] : s;
    if (in_range && !live) {   // padded row: the reference leaves its -1 fill (grid.cu:422-423)
      const size_t o = ((size_t)n * P1 + s) * K;
      for (int k = 0; k < K; ++k) { dists[o + k] = -1.f; idxs[o + k] = (IdxT)-1; }
    }

    int cnt = 0;
    bool failed = false;
    if (live) {
      const float delta = prm[D];
      const float inv_delta = 1.0f / delta;
      int lo[D], hi[D], res[D];
      float qc[D];
      bool nonempty = true;
#pragma unroll
      for (int d = 0; d < D; ++d) {
        const float rel = __fsub_rn(q[d], prm[d]);
        res[d] = (int)prm[D + 1 + d];
        lo[d] = max(__float2int_rd(__fmul_rn(__fsub_rn(rel, r), delta)), 0);       // grid.cu:305-316
        hi[d] = min(__float2int_rd(__fmul_rn(__fadd_rn(rel, r), delta)), res[d] - 1);
        nonempty = nonempty && (lo[d] <= hi[d]);
        qc[d] = rel * delta;
      }
      const int grid_total = (int)prm[2 * D + 1];
      if (nonempty) {
        const int zlo = lo[D - 1], zhi = hi[D - 1];
        const int ylo = (D == 3) ? lo[1] : 0, yhi = (D == 3) ? hi[1] : 0;
        // ---- trial radius from the block's point density ----
        // (A three-scale estimate -- own cell / 3^D cells / block, log-log interpolation of the box that holds
        // 1.9 lambda points -- was measured too: it helps clumped clouds (C2's iso-points: 0.58 -> 0.40 ms) but its
        // single-cell count is noisy on uniform data (C3: 0.56 -> 0.76 ms), and on C2 the pruned group kernel
        // stays ahead of both.)
        int nblock = 0;
        for (int x = lo[0]; x <= hi[0]; ++x)
          for (int y = ylo; y <= yhi; ++y) {
            const int c0 = (D == 3) ? (x * res[1] + y) * res[2] : x * res[1];
            const int a = off2[c0 + zlo];
            const int b = (c0 + zhi + 1 == grid_total) ? len2 : off2[c0 + zhi + 1];
            nblock += b - a;
          }
        // (a block that fits the column is simply collected whole: exact without any estimate)
        float tau = r2;
        if (nblock > COLLECT_CAP) {
          float cells = (float)(hi[0] - lo[0] + 1) * (float)(zhi - zlo + 1);
          if (D == 3) cells *= (float)(yhi - ylo + 1);
          // radius (in cells) of the ball expected to hold lambda points at the block's mean density
          const float vol = lambda * cells / (float)nblock;                  // in cell volumes
          float rc = (D == 3) ? cbrtf(vol * 0.238732415f) : sqrtf(vol * 0.318309886f);
          // a ball that sticks out of the grid holds fewer points: grow it by the part cut off (per axis the
          // fraction of the diameter inside, one fixed-point step) -- the estimate only has to be roughly right
          float inside = 1.0f;
#pragma unroll
          for (int d = 0; d < D; ++d) {
            const float t = fminf(qc[d], (float)res[d] - qc[d]);             // cells to the nearer grid face
            inside *= fminf(fmaxf(0.5f * (t / rc + 1.0f), 0.5f), 1.0f);
          }
          rc = (D == 3) ? rc * cbrtf(1.0f / inside) : rc * sqrtf(1.0f / inside);
          const float re = rc * inv_delta;
          tau = fminf(r2, re * re);
        }
        // Up to three passes: the trial radius assumes the block's MEAN density and a volume-filling cloud; on a
        // surface (points ~ r^2) or in a cluster the ball overflows the column, next to a void it comes up short.
        // A failed pass re-scales the radius from the counts it saw, and only a query that is still off after
        // the third goes to the warp-cooperative search.
        bool exact = false;
        for (int attempt = 0; attempt < 3 && !exact; ++attempt) {
          cnt = 0;
          int cntq = 0;                        // candidates inside a quarter of the trial tau (half the radius)
          const float tauq = 0.25f * tau;
          // ---- one pass over the cells the trial ball touches ----
          // (a per-thread run cursor inside one flat candidate loop was tried: the cursor code then runs with ~5
          // live lanes per instruction and costs more than the per-run trip-count divergence of these nested loops)
          for (int x = lo[0]; x <= hi[0]; ++x) {
            const float gx2 = axis_gap2(qc[0], x, inv_delta);
            if (gx2 > tau) continue;
            for (int y = ylo; y <= yhi; ++y) {
              float gxy2 = gx2;
              int c0;
              if (D == 3) {
                gxy2 += axis_gap2(qc[1], y, inv_delta);
                if (gxy2 > tau) continue;
                c0 = (x * res[1] + y) * res[2];
              } else {
                c0 = x * res[1];
              }
              const float gz = sqrtf(fmaxf(tau - gxy2, 0.0f)) * delta + 2e-3f;
              const int za = max(zlo, __float2int_ru(qc[D - 1] - 1.0f - gz));
              const int zb = min(zhi, __float2int_rd(qc[D - 1] + gz));
              if (za > zb) continue;
              const int start = off2[c0 + za];
              const int end = (c0 + zb + 1 == grid_total) ? len2 : off2[c0 + zb + 1];
              // four candidates per step: their 4 D loads are in flight together (the loop is latency bound:
              // one query per thread leaves 16 warps per SM)
              for (int j = start; j < end; j += 4) {
                float dd[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                  const int jj = min(j + u, end - 1);
                  dd[u] = sqdist_ref<D>(pts2 + (size_t)jj * D, q);
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                  if (j + u < end && dd[u] <= tau) {
                    if (cnt < COLLECT_CAP)
                      s_keys[cnt * COLLECT_THREADS + tid] =
                          ((unsigned long long)__float_as_uint(dd[u]) << 32) | (unsigned)sid2[j + u];
                    ++cnt;
                    cntq += dd[u] <= tauq;
                  }
                }
              }
            }
          }
          // exact unless the ball overflowed the column, or was cut short of K by a trial radius below r.
          // The failed pass measured the count at two radii (tau and tau / 4): count ~ tau^e locally, with e =
          // 1.5 in a volume, 1 on a surface and far below that inside a clump of near-coincident points (iso-points
          // of a random-init SIREN: e ~ 0.4) -- the next tau is read off that power law.
          if (cnt > COLLECT_CAP || (cnt < K && tau < r2)) {
            float e = 0.5f * __log2f((float)max(cnt, 1) / fmaxf((float)cntq, 0.5f));
            e = fminf(fmaxf(e, 0.2f), 1.5f);
            const float target = cnt > COLLECT_CAP ? lambda : 1.5f * lambda;
            const float scale = exp2f(__log2f(target / (float)max(cnt, 1)) / e);
            tau = fminf(r2, tau * fminf(fmaxf(scale, 1e-4f), 64.0f));
          } else {
            exact = true;
          }
        }
        failed = !exact;
      }
      if (!failed) {
        // ---- emit the K smallest keys in ascending order (keys are unique: the index is part of them), two per
        //      pass over the column: the smallest and second smallest key above the last one written ----
        const size_t o = ((size_t)n * P1 + row) * K;
        unsigned long long prev = 0;
        bool have_prev = false;
        for (int k = 0; k < K; k += 2) {
          unsigned long long b1 = ~0ull, b2 = ~0ull;
          if (k < cnt) {
            for (int i = 0; i < cnt; ++i) {
              const unsigned long long key = s_keys[i * COLLECT_THREADS + tid];
              if (!have_prev || key > prev) {
                const bool lt1 = key < b1, lt2 = key < b2;
                b2 = lt1 ? b1 : (lt2 ? key : b2);
                b1 = lt1 ? key : b1;
              }
            }
          }
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const unsigned long long best = u ? b2 : b1;
            if (k + u >= K) break;
            if (best != ~0ull) {
              dists[o + k + u] = __uint_as_float((unsigned)(best >> 32));
              idxs[o + k + u] = (IdxT)(int)(unsigned)(best & 0xffffffffu);
              prev = best;
              have_prev = true;
            } else {
              dists[o + k + u] = -1.f;
              idxs[o + k + u] = (IdxT)-1;
            }
          }
        }
      }
    }
    // ---- the warp takes over the queries whose trial ball failed ----
    unsigned fm = __ballot_sync(0xffffffffu, failed);
    while (fm) {
      const int src = __ffs(fm) - 1;
      fm &= fm - 1;
      float qq[D];
#pragma unroll
      for (int d = 0; d < D; ++d) qq[d] = __shfl_sync(0xffffffffu, q[d], src);
      const int nn = __shfl_sync(0xffffffffu, n, src);
      const int rr = __shfl_sync(0xffffffffu, row, src);
      const int l2 = lengths2 ? (int)min((long long)lengths2[nn], (long long)P2) : P2;
      const float r_ = rs[nn];
      float bd;
      int bi;
      warp_pruned_search<D>(qq, params + (size_t)nn * PS, __fmul_rn(r_, r_), r_,
                            sorted_points2 + (size_t)nn * P2 * D, cell_off2 + (size_t)nn * G,
                            sorted_idxs2 + (size_t)nn * P2, l2, K, bd, bi);
      if (lane < K) {
        const size_t o = ((size_t)nn * P1 + rr) * K + lane;
        const bool found = bi != INT_MAX;
        dists[o] = found ? bd : -1.f;
        idxs[o] = found ? (IdxT)bi : (IdxT)-1;
      }
    }
  }
}

// x (N,M,U), idxs (N,L,K) -> out (N,L,K,U); 0 where idx < 0   (frnn.py:304-352)
template <typename IdxT>
__global__ void __launch_bounds__(256)
frnn_gather_kernel(const float* __restrict__ x, const IdxT* __restrict__ idxs, int N, int M, int L,
                   int K, int U, float* __restrict__ out) {
  const long long total = (long long)N * L * K * U;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int u = (int)(i % U);
    const long long e = i / U;  // (n,l,k) flat
    const int n = (int)(e / ((long long)L * K));
    const long long j = (long long)idxs[e];
    out[i] = (j >= 0) ? x[((size_t)n * M + j) * U + u] : 0.f;
  }
}

// grad_out (N,L,K,U) scattered back into grad_x (N,M,U) (+=), skipping idx < 0
template <typename IdxT>
__global__ void __launch_bounds__(256)
frnn_gather_bwd_kernel(const float* __restrict__ grad_out, const IdxT* __restrict__ idxs, int N,
                       int M, int L, int K, int U, float* __restrict__ grad_x) {
  const long long total = (long long)N * L * K * U;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int u = (int)(i % U);
    const long long e = i / U;
    const int n = (int)(e / ((long long)L * K));
    const long long j = (long long)idxs[e];
    if (j >= 0) atomicAdd(&grad_x[((size_t)n * M + j) * U + u], grad_out[i]);
  }
}

// d dists / d points (backward.cu:8-74): one warp-coalesced pass over (n,p1,k); the p1 side is
// reduced over k in registers before a single atomic per coordinate.
template <int D>
__global__ void __launch_bounds__(256)
frnn_backward_kernel(const float* __restrict__ points1, const float* __restrict__ points2,
                     const int64_t* __restrict__ lengths1, const int64_t* __restrict__ lengths2,
                     const int64_t* __restrict__ idxs, const float* __restrict__ grad_dists, int N,
                     int P1, int P2, int K, float* __restrict__ grad_points1,
                     float* __restrict__ grad_points2) {
  const long long total = (long long)N * P1;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(i / P1), p1 = (int)(i % P1);
    if (p1 >= lengths1[n]) continue;
    const long long num2 = lengths2[n];
    float a[D], g1[D];
#pragma unroll
    for (int d = 0; d < D; ++d) { a[d] = points1[i * D + d]; g1[d] = 0.f; }
    for (int k = 0; k < K && k < num2; ++k) {
      const long long j = idxs[i * K + k];
      if (j < 0) continue;
      const float g = grad_dists[i * K + k];
#pragma unroll
      for (int d = 0; d < D; ++d) {
        const float diff = 2.0f * g * (a[d] - points2[((size_t)n * P2 + j) * D + d]);
        g1[d] += diff;
        atomicAdd(&grad_points2[((size_t)n * P2 + j) * D + d], -diff);
      }
    }
#pragma unroll
    for (int d = 0; d < D; ++d) atomicAdd(&grad_points1[i * D + d], g1[d]);
  }
}

template <int D, typename IdxT>
static int launch_query(bool exhaustive, int gw, int blocks, cudaStream_t st, const float* qp, const int* qo,
                        const int64_t* l1, const int64_t* l2, const float* sp2, const int* off2,
                        const int* sid2, const float* params, const float* rs, int N, int P1, int P2,
                        int G, int K, float* dists, IdxT* idxs) {
  if (exhaustive) {
    if (gw == 8) frnn_query_exhaustive_kernel<D, 8, IdxT><<<blocks, 256, 0, st>>>(qp, qo, l1, l2, sp2, off2, sid2, params, rs, N, P1, P2, G, K, dists, idxs);
    else if (gw == 16) frnn_query_exhaustive_kernel<D, 16, IdxT><<<blocks, 256, 0, st>>>(qp, qo, l1, l2, sp2, off2, sid2, params, rs, N, P1, P2, G, K, dists, idxs);
    else frnn_query_exhaustive_kernel<D, 32, IdxT><<<blocks, 256, 0, st>>>(qp, qo, l1, l2, sp2, off2, sid2, params, rs, N, P1, P2, G, K, dists, idxs);
    ISO_CHECK_LAUNCH("frnn_query_exhaustive_kernel");
    return ISOB200_OK;
  }
  if (gw == 8)
    frnn_query_kernel<D, 8, IdxT><<<blocks, 256, 0, st>>>(qp, qo, l1, l2, sp2, off2, sid2, params, rs, N, P1, P2, G, K, dists, idxs);
  else if (gw == 16)
    frnn_query_kernel<D, 16, IdxT><<<blocks, 256, 0, st>>>(qp, qo, l1, l2, sp2, off2, sid2, params, rs, N, P1, P2, G, K, dists, idxs);
  else
    frnn_query_kernel<D, 32, IdxT><<<blocks, 256, 0, st>>>(qp, qo, l1, l2, sp2, off2, sid2, params, rs, N, P1, P2, G, K, dists, idxs);
  ISO_CHECK_LAUNCH("frnn_query_kernel");
  return ISOB200_OK;
}

}  // namespace isob200

using namespace isob200;

extern "C" {

// == frnn._C.find_nbrs_cuda (grid.cu:384-440), minus the allocation: dists (N,P1,K) f32 and
// idxs (N,P1,K) are caller tensors and are fully written (incl. -1 padding).
//   q_points : query coordinates in processing order (pass the cell-sorted copy for locality)
//   q_order  : processing slot -> original row (sorted_points1_idxs), or NULL for identity
//   idx_is_i64: 1 -> idxs is int64 (reference API), 0 -> int32 (internal fused consumers)
//   group_width: 0 = auto (smallest of 8/16/32 that is >= K), optionally OR-ed with a traversal mode in
//                bits 8..9: 0 = auto (exhaustive when the grid holds < 1 point per cell on average; else
//                thread-per-query collect-then-select for K <= 20 and the pruned best-first group search
//                above that), 1 = exhaustive, 2 = pruned, 3 = collect.  Results are identical.
int isob200_frnn_find_nbrs(const float* q_points, const int* q_order, const int64_t* lengths1,
                           const int64_t* lengths2, const float* sorted_points2, const int* cell_off2,
                           const int* sorted_idxs2, const float* params, const float* rs, int N,
                           int P1, int P2, int D, int G, int K, float* dists, void* idxs,
                           int idx_is_i64, int group_width, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  ISO_CHECK_ARG(D == 2 || D == 3, "for now only 2D and 3D are supported");
  ISO_CHECK_ARG(K >= 1 && K <= 32, "Invalid range: K=%d must be in [1, 32]", K);
  ISO_CHECK_ARG(N >= 0 && P1 >= 0 && P2 >= 0 && G > 0, "find_nbrs: bad sizes");
  if ((long long)N * P1 == 0) return ISOB200_OK;
  ISO_CHECK_ARG(q_points && sorted_points2 && cell_off2 && sorted_idxs2 && params && rs && dists && idxs,
                "find_nbrs: null pointer");
  const int mode = (group_width >> 8) & 3;
  int gw = group_width & 0xff;
  if (gw == 0) gw = (K <= 8) ? 8 : (K <= 16 ? 16 : 32);
  // bit 10: the caller knows that the radius spans many point spacings (the K-th neighbour lies far inside r):
  // on a sparse grid the pruned traversal then beats the exhaustive one (C2's resample tree: 0.29 vs 0.35 ms)
  const bool far_radius = (group_width >> 10) & 1;
  const bool sparse = (long long)P2 < (long long)G;
  const bool exhaustive = mode == 1 || (mode == 0 && sparse && !far_radius);
  // dense grid and a K whose trial ball fits a shared-memory column: thread-per-query collect-then-select.
  // (Measured on the sparse grid of BASELINE config 2 -- iso-surface points, ~0.1 per cell, K = 9: 0.65 ms against
  // 0.36 ms for the exhaustive group kernel; there the candidate block often exceeds the column and the volume-based
  // trial radius needs its second pass.)
  const bool collect = mode == 3 || (mode == 0 && !sparse && K <= 20);
  if (collect) {
    ISO_CHECK_ARG(K <= 32, "find_nbrs: collect mode needs K <= 32");
    const float lambda = fminf((float)K + 4.0f * sqrtf((float)K), (float)(COLLECT_CAP - 8));
    const size_t smem = (size_t)COLLECT_CAP * COLLECT_THREADS * sizeof(unsigned long long);
    const long long items_c = (long long)N * P1;
    long long need_c = (items_c + COLLECT_THREADS - 1) / COLLECT_THREADS;
    const long long cap_c = (long long)kNumSMs * 4 * 32;
    const int blocks_c = (int)(need_c < cap_c ? need_c : cap_c);
#define QC(DD, T)                                                                                          \
  frnn_query_collect_kernel<DD, T><<<blocks_c, COLLECT_THREADS, smem, st>>>(                               \
      q_points, q_order, lengths1, lengths2, sorted_points2, cell_off2, sorted_idxs2, params, rs, N, P1,  \
      P2, G, K, lambda, dists, (T*)idxs)
    if (D == 3) { if (idx_is_i64) QC(3, int64_t); else QC(3, int); }
    else { if (idx_is_i64) QC(2, int64_t); else QC(2, int); }
#undef QC
    ISO_CHECK_LAUNCH("frnn_query_collect_kernel");
    return ISOB200_OK;
  }
  ISO_CHECK_ARG((gw == 8 || gw == 16 || gw == 32) && gw >= K, "find_nbrs: group_width %d invalid for K=%d", gw, K);
  const long long items = (long long)N * P1;
  const int groups = 256 / gw;
  long long need = (items + groups - 1) / groups;
  const long long cap = (long long)kNumSMs * 8 * 16;  // 8 resident CTAs/SM, 16 waves max
  const int blocks = (int)(need < cap ? need : cap);
#define Q(DD, T) launch_query<DD, T>(exhaustive, gw, blocks, st, q_points, q_order, lengths1, lengths2, sorted_points2, cell_off2, sorted_idxs2, params, rs, N, P1, P2, G, K, dists, (T*)idxs)
  if (D == 3) return idx_is_i64 ? Q(3, int64_t) : Q(3, int);
  return idx_is_i64 ? Q(2, int64_t) : Q(2, int);
#undef Q
}

int isob200_frnn_gather(const float* x, const void* idxs, int idx_is_i64, int N, int M, int L, int K,
                        int U, float* out, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  const long long total = (long long)N * L * K * U;
  if (total == 0) return ISOB200_OK;
  ISO_CHECK_ARG(x && idxs && out, "frnn_gather: null pointer");
  const int g = grid_for(total, 256, 8);
  if (idx_is_i64) frnn_gather_kernel<int64_t><<<g, 256, 0, st>>>(x, (const int64_t*)idxs, N, M, L, K, U, out);
  else frnn_gather_kernel<int><<<g, 256, 0, st>>>(x, (const int*)idxs, N, M, L, K, U, out);
  ISO_CHECK_LAUNCH("frnn_gather_kernel");
  return ISOB200_OK;
}

int isob200_frnn_gather_backward(const float* grad_out, const void* idxs, int idx_is_i64, int N, int M,
                                 int L, int K, int U, float* grad_x, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  const long long total = (long long)N * L * K * U;
  if (total == 0) return ISOB200_OK;
  ISO_CHECK_ARG(grad_out && idxs && grad_x, "frnn_gather_backward: null pointer");
  const int g = grid_for(total, 256, 8);
  if (idx_is_i64) frnn_gather_bwd_kernel<int64_t><<<g, 256, 0, st>>>(grad_out, (const int64_t*)idxs, N, M, L, K, U, grad_x);
  else frnn_gather_bwd_kernel<int><<<g, 256, 0, st>>>(grad_out, (const int*)idxs, N, M, L, K, U, grad_x);
  ISO_CHECK_LAUNCH("frnn_gather_bwd_kernel");
  return ISOB200_OK;
}

// == frnn._C.frnn_backward_cuda (backward.cu:76-147); grad_points{1,2} must be zero-initialised.
int isob200_frnn_backward(const float* points1, const float* points2, const int64_t* lengths1,
                          const int64_t* lengths2, const int64_t* idxs, const float* grad_dists, int N,
                          int P1, int P2, int D, int K, float* grad_points1, float* grad_points2,
                          void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  ISO_CHECK_ARG(D == 2 || D == 3, "for now only 2D and 3D are supported");
  if ((long long)N * P1 * K == 0) return ISOB200_OK;
  ISO_CHECK_ARG(points1 && points2 && lengths1 && lengths2 && idxs && grad_dists && grad_points1 && grad_points2,
                "frnn_backward: null pointer");
  const int g = grid_for((long long)N * P1, 256, 8);
  if (D == 3) frnn_backward_kernel<3><<<g, 256, 0, st>>>(points1, points2, lengths1, lengths2, idxs, grad_dists, N, P1, P2, K, grad_points1, grad_points2);
  else frnn_backward_kernel<2><<<g, 256, 0, st>>>(points1, points2, lengths1, lengths2, idxs, grad_dists, N, P1, P2, K, grad_points1, grad_points2);
  ISO_CHECK_LAUNCH("frnn_backward_kernel");
  return ISOB200_OK;
}

}  // extern "C"
