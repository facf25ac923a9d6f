// Level-set projection (Newton iteration) kernels for sm_100a.
//
// Replaces the per-iteration PyTorch op chain of UniformProjection._project_points
// (DSS/models/levelset_sampling.py:313-342): boolean-mask gathers/scatters (`x[mask]`,
// `x[mask] = y` -> nonzero + index kernels with a host sync each), eps_denom / normalize /
// clamp_max elementwise kernels, and the `(~not_converged).all()` read-back.
//
// One fused kernel per Newton iteration:
//   normals[act[i]] = grad[i]                                  (:322)
//   still = |sdf[i]| > tol ; not_converged[act[i]] = still     (:326-328)
//   if still and this is not the last evaluation:              (:333-342)
//       move = sdf * grad / eps_denom(|grad|^2, 1e-17)
//       move = normalize(move, eps=1e-15) * min(|move|, 0.1)
//       points[act[i]] -= move
//   still-active indices are compacted IN ORDER into the next active list with a warp-ballot
//   + single-pass decoupled-look-back scan (no host sync, no second pass); the new count
//   lands in a device counter the host reads once per iteration only because the opaque
//   nn.Module SDF callback needs a tensor shape.
// HBM traffic per active point-iteration: read idx 4 + sdf 4 + grad 12 + xyz 12, write
// normal 12 + xyz 12 + flag 1 + idx 4  (= 61 B; SURVEY 8d quotes 56 B without idx/flag).
#include "common.cuh"
#include <math.h>

namespace isob200 {

constexpr int PJ_THREADS = 256;
constexpr int PJ_ITEMS = 4;
constexpr int PJ_TILE = PJ_THREADS * PJ_ITEMS;

// tile status word for the decoupled look-back: [31:30] flag, [29:0] value
constexpr unsigned LB_FLAG_AGG = 1u << 30;   // tile aggregate available
constexpr unsigned LB_FLAG_INC = 2u << 30;   // inclusive prefix available
constexpr unsigned LB_VALUE_MASK = (1u << 30) - 1u;

__device__ __forceinline__ unsigned ld_volatile_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}

// eps_denom(x, eps) of DSS/utils/mathHelper.py:14-18: (sign(x) + [x == 0]) * max(|x|, eps)
__device__ __forceinline__ float eps_denom_f(float x, float eps) {
  const float s = (x > 0.f) ? 1.f : ((x < 0.f) ? -1.f : 1.f);
  return __fmul_rn(s, fmaxf(fabsf(x), eps));
}

// ws layout: [0] ticket counter, [1] unused, [2 .. 2+tiles) tile status words.  Zeroed by the
// launcher (cudaMemsetAsync) before every launch.
__global__ void __launch_bounds__(PJ_THREADS)
project_step_kernel(float* __restrict__ points, float* __restrict__ normals,
                    unsigned char* __restrict__ not_converged, const int* __restrict__ act_in,
                    int A, const int* __restrict__ a_dev, const float* __restrict__ sdf,
                    const float* __restrict__ grad,
                    float tol, float max_step, int do_update, int* __restrict__ act_out,
                    float* __restrict__ next_points, int* __restrict__ count_out,
                    unsigned* __restrict__ ws) {
  __shared__ int s_tile;
  __shared__ int s_warp[PJ_THREADS / 32];
  __shared__ int s_prefix;
  if (threadIdx.x == 0) s_tile = (int)atomicAdd(&ws[0], 1u);  // ticket => forward progress
  __syncthreads();
  const int tile = s_tile;
  if (a_dev) {  // live row count kept on the device (sync-free loop): A is only the upper bound
    const int a = *a_dev;
    A = a < A ? a : A;
    if ((long long)tile * PJ_TILE >= A) {  // later tickets are all past the end too: nobody looks back here
      if (tile == 0 && threadIdx.x == 0) *count_out = 0;
      return;
    }
  }
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  // blocked assignment: thread t owns items [base + t*ITEMS, +ITEMS) => order-preserving ranks
  const int base = tile * PJ_TILE + threadIdx.x * PJ_ITEMS;

  int keep[PJ_ITEMS];
  int id[PJ_ITEMS];
  float nx[PJ_ITEMS][3];   // position after the update (the next SDF evaluation's input row)
  int cnt = 0;
#pragma unroll
  for (int j = 0; j < PJ_ITEMS; ++j) {
    const int i = base + j;
    keep[j] = 0;
    id[j] = -1;
    if (i < A) {
      const int p = act_in ? act_in[i] : i;
      id[j] = p;
      const float f = sdf[i];
      const float gx = grad[3 * i + 0], gy = grad[3 * i + 1], gz = grad[3 * i + 2];
      normals[3 * (size_t)p + 0] = gx;
      normals[3 * (size_t)p + 1] = gy;
      normals[3 * (size_t)p + 2] = gz;
      const bool still = fabsf(f) > tol;
      if (not_converged) not_converged[p] = still ? 1 : 0;
      if (still) {
        keep[j] = 1;
        {
          const float* q0 = points + 3 * (size_t)p;
          nx[j][0] = q0[0]; nx[j][1] = q0[1]; nx[j][2] = q0[2];
        }
        if (do_update) {
          const float ss = __fadd_rn(__fadd_rn(__fmul_rn(gx, gx), __fmul_rn(gy, gy)), __fmul_rn(gz, gz));
          const float den = eps_denom_f(ss, 1.0e-17f);
          const float mx = __fmul_rn(f, __fdiv_rn(gx, den));
          const float my = __fmul_rn(f, __fdiv_rn(gy, den));
          const float mz = __fmul_rn(f, __fdiv_rn(gz, den));
          const float nrm = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(mx, mx), __fmul_rn(my, my)), __fmul_rn(mz, mz)));
          const float dn = fmaxf(nrm, 1e-15f);        // F.normalize(eps=1e-15)
          const float len = fminf(nrm, max_step);     // clamp_max(0.1)
          float* q = points + 3 * (size_t)p;
          nx[j][0] = __fsub_rn(nx[j][0], __fmul_rn(__fdiv_rn(mx, dn), len));
          nx[j][1] = __fsub_rn(nx[j][1], __fmul_rn(__fdiv_rn(my, dn), len));
          nx[j][2] = __fsub_rn(nx[j][2], __fmul_rn(__fdiv_rn(mz, dn), len));
          q[0] = nx[j][0]; q[1] = nx[j][1]; q[2] = nx[j][2];
        }
      }
    }
    cnt += keep[j];
  }
  // block exclusive scan of per-thread counts
  int incl = cnt;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) s_warp[w] = incl;
  __syncthreads();
  int woff = 0, total = 0;
#pragma unroll
  for (int i = 0; i < PJ_THREADS / 32; ++i) {
    const int t = s_warp[i];
    if (i < w) woff += t;
    total += t;
  }
  // decoupled look-back (one thread; tiles are few: A / 1024)
  if (threadIdx.x == 0) {
    unsigned* status = ws + 2;
    int prefix = 0;
    if (tile == 0) {
      __threadfence();
      atomicExch(&status[0], LB_FLAG_INC | (unsigned)total);
    } else {
      atomicExch(&status[tile], LB_FLAG_AGG | (unsigned)total);
      int t = tile - 1;
      while (true) {
        const unsigned s = ld_volatile_u32(&status[t]);
        if ((s >> 30) == 0) continue;  // not published yet
        prefix += (int)(s & LB_VALUE_MASK);
        if (s & LB_FLAG_INC) break;
        --t;
      }
      atomicExch(&status[tile], LB_FLAG_INC | (unsigned)(prefix + total));
    }
    s_prefix = prefix;
    if ((long long)(tile + 1) * PJ_TILE >= A) *count_out = prefix + total;  // last tile
  }
  __syncthreads();
  int pos = s_prefix + woff + incl - cnt;
#pragma unroll
  for (int j = 0; j < PJ_ITEMS; ++j)
    if (keep[j]) {
      act_out[pos] = id[j];
      if (next_points) {   // compacted input of the next SDF evaluation: no separate gather pass
        next_points[3 * (size_t)pos + 0] = nx[j][0];
        next_points[3 * (size_t)pos + 1] = nx[j][1];
        next_points[3 * (size_t)pos + 2] = nx[j][2];
      }
      ++pos;
    }
}

// One iteration of SphereTracing.project_points (levelset_sampling.py:733-786) on the active rays:
//   eval[act[i]] = sdf[i] ; grad_out[act[i]] = grad[i]                           (:742-758)
//   still = |sdf[i]| > 0.1 tol                   (active rays are inside the sphere by construction, :760-761)
//   if still and this is not the last evaluation:                                (:764-776)
//       move = alpha * sdf * dir ; move = normalize(move, eps=1e-15) * min(|move|, 0.1)
//       p' = p + move ; inside = |p'| < padding + radius
//       inside: points[act[i]] = p', the ray stays active; else the ray keeps its old position and retires
// Same ticket + decoupled look-back compaction as project_step_kernel (order-preserving).
__global__ void __launch_bounds__(PJ_THREADS)
trace_step_kernel(float* __restrict__ points, const float* __restrict__ dirs, float* __restrict__ eval,
                  float* __restrict__ grad_out, const int* __restrict__ act_in, int A,
                  const int* __restrict__ a_dev, const float* __restrict__ sdf, const float* __restrict__ grad,
                  float active_tol, float alpha, float max_step, float bound, int do_update,
                  int* __restrict__ act_out, float* __restrict__ next_points, int* __restrict__ count_out,
                  unsigned* __restrict__ ws) {
  __shared__ int s_tile;
  __shared__ int s_warp[PJ_THREADS / 32];
  __shared__ int s_prefix;
  if (threadIdx.x == 0) s_tile = (int)atomicAdd(&ws[0], 1u);
  __syncthreads();
  const int tile = s_tile;
  if (a_dev) {
    const int a = *a_dev;
    A = a < A ? a : A;
    if ((long long)tile * PJ_TILE >= A) {
      if (tile == 0 && threadIdx.x == 0) *count_out = 0;
      return;
    }
  }
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int base = tile * PJ_TILE + threadIdx.x * PJ_ITEMS;
  int keep[PJ_ITEMS];
  int id[PJ_ITEMS];
  float nx[PJ_ITEMS][3];
  int cnt = 0;
#pragma unroll
  for (int j = 0; j < PJ_ITEMS; ++j) {
    const int i = base + j;
    keep[j] = 0;
    id[j] = -1;
    if (i < A) {
      const int p = act_in ? act_in[i] : i;
      id[j] = p;
      const float f = sdf[i];
      eval[p] = f;
      if (grad_out) {
        grad_out[3 * (size_t)p + 0] = grad[3 * i + 0];
        grad_out[3 * (size_t)p + 1] = grad[3 * i + 1];
        grad_out[3 * (size_t)p + 2] = grad[3 * i + 2];
      }
      if (fabsf(f) > active_tol) {
        float* q = points + 3 * (size_t)p;
        nx[j][0] = q[0]; nx[j][1] = q[1]; nx[j][2] = q[2];
        keep[j] = 1;
        if (do_update) {
          const float af = __fmul_rn(alpha, f);
          const float mx = __fmul_rn(af, dirs[3 * (size_t)p + 0]);
          const float my = __fmul_rn(af, dirs[3 * (size_t)p + 1]);
          const float mz = __fmul_rn(af, dirs[3 * (size_t)p + 2]);
          const float nrm = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(mx, mx), __fmul_rn(my, my)), __fmul_rn(mz, mz)));
          const float dn = fmaxf(nrm, 1e-15f);        // F.normalize(eps=1e-15)
          const float len = fminf(nrm, max_step);     // clamp_max(0.1)
          nx[j][0] = __fadd_rn(nx[j][0], __fmul_rn(__fdiv_rn(mx, dn), len));
          nx[j][1] = __fadd_rn(nx[j][1], __fmul_rn(__fdiv_rn(my, dn), len));
          nx[j][2] = __fadd_rn(nx[j][2], __fmul_rn(__fdiv_rn(mz, dn), len));
          const float r = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(nx[j][0], nx[j][0]), __fmul_rn(nx[j][1], nx[j][1])),
                                          __fmul_rn(nx[j][2], nx[j][2])));
          if (r < bound) { q[0] = nx[j][0]; q[1] = nx[j][1]; q[2] = nx[j][2]; }
          else keep[j] = 0;                            // left the sphere: old position stays, ray retires
        }
      }
    }
    cnt += keep[j];
  }
  int incl = cnt;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) s_warp[w] = incl;
  __syncthreads();
  int woff = 0, total = 0;
#pragma unroll
  for (int i = 0; i < PJ_THREADS / 32; ++i) {
    const int t = s_warp[i];
    if (i < w) woff += t;
    total += t;
  }
  if (threadIdx.x == 0) {
    unsigned* status = ws + 2;
    int prefix = 0;
    if (tile == 0) {
      __threadfence();
      atomicExch(&status[0], LB_FLAG_INC | (unsigned)total);
    } else {
      atomicExch(&status[tile], LB_FLAG_AGG | (unsigned)total);
      int t = tile - 1;
      while (true) {
        const unsigned s = ld_volatile_u32(&status[t]);
        if ((s >> 30) == 0) continue;
        prefix += (int)(s & LB_VALUE_MASK);
        if (s & LB_FLAG_INC) break;
        --t;
      }
      atomicExch(&status[tile], LB_FLAG_INC | (unsigned)(prefix + total));
    }
    s_prefix = prefix;
    if ((long long)(tile + 1) * PJ_TILE >= A) *count_out = prefix + total;
  }
  __syncthreads();
  int pos = s_prefix + woff + incl - cnt;
#pragma unroll
  for (int j = 0; j < PJ_ITEMS; ++j)
    if (keep[j]) {
      act_out[pos] = id[j];
      if (next_points) {
        next_points[3 * (size_t)pos + 0] = nx[j][0];
        next_points[3 * (size_t)pos + 1] = nx[j][1];
        next_points[3 * (size_t)pos + 2] = nx[j][2];
      }
      ++pos;
    }
}

// Order-preserving compaction of the valid (converged) rows of (points, normals): _filter_projection_result
// (levelset_sampling.py:59-65 -> DSS/utils/__init__.py:149-169) for one packed cloud, in one pass with
// the same ticket + decoupled look-back scan as project_step_kernel.  count_out receives the number
// of survivors (the only value the host reads back: it is the output shape).
__global__ void __launch_bounds__(PJ_THREADS)
compact_valid_kernel(const float* __restrict__ points, const float* __restrict__ normals,
                     const unsigned char* __restrict__ valid, int M, float* __restrict__ out_points,
                     float* __restrict__ out_normals, int* __restrict__ count_out, unsigned* __restrict__ ws) {
  __shared__ int s_tile;
  __shared__ int s_warp[PJ_THREADS / 32];
  __shared__ int s_prefix;
  if (threadIdx.x == 0) s_tile = (int)atomicAdd(&ws[0], 1u);
  __syncthreads();
  const int tile = s_tile;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int base = tile * PJ_TILE + threadIdx.x * PJ_ITEMS;
  int keep[PJ_ITEMS];
  int cnt = 0;
#pragma unroll
  for (int j = 0; j < PJ_ITEMS; ++j) {
    const int i = base + j;
    keep[j] = (i < M) && valid[i];
    cnt += keep[j];
  }
  int incl = cnt;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) s_warp[w] = incl;
  __syncthreads();
  int woff = 0, total = 0;
#pragma unroll
  for (int i = 0; i < PJ_THREADS / 32; ++i) {
    const int t = s_warp[i];
    if (i < w) woff += t;
    total += t;
  }
  if (threadIdx.x == 0) {
    unsigned* status = ws + 2;
    int prefix = 0;
    if (tile == 0) {
      __threadfence();
      atomicExch(&status[0], LB_FLAG_INC | (unsigned)total);
    } else {
      atomicExch(&status[tile], LB_FLAG_AGG | (unsigned)total);
      int t = tile - 1;
      while (true) {
        const unsigned s = ld_volatile_u32(&status[t]);
        if ((s >> 30) == 0) continue;
        prefix += (int)(s & LB_VALUE_MASK);
        if (s & LB_FLAG_INC) break;
        --t;
      }
      atomicExch(&status[tile], LB_FLAG_INC | (unsigned)(prefix + total));
    }
    s_prefix = prefix;
    if ((long long)(tile + 1) * PJ_TILE >= M) *count_out = prefix + total;
  }
  __syncthreads();
  int pos = s_prefix + woff + incl - cnt;
#pragma unroll
  for (int j = 0; j < PJ_ITEMS; ++j)
    if (keep[j]) {
      const size_t i = (size_t)(base + j);
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        out_points[3 * (size_t)pos + c] = points[3 * i + c];
        if (normals) out_normals[3 * (size_t)pos + c] = normals[3 * i + c];
      }
      ++pos;
    }
}

// dst[i, :] = src[idx[i], :]   (curr_points = points_packed[not_converged], :315)
__global__ void __launch_bounds__(256)
gather_rows3_kernel(const float* __restrict__ src, const int* __restrict__ idx, int A,
                    float* __restrict__ dst) {
  // 3 consecutive threads move one row => dst writes fully coalesced
  const long long total = 3ll * A;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / 3), c = (int)(i - 3ll * r);
    dst[i] = src[3 * (size_t)idx[r] + c];
  }
}

// ---------------------------------------------------------------------------------------
// Fully fused projection for a built-in analytic SDF (unit sphere |x| - R): all iterations in
// registers, one read + one write per point (37 B/point).  Gradient as autograd
// gives it for x.norm(dim=-1): x / |x| (0 at the origin).
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
project_sphere_kernel(float* __restrict__ points, float* __restrict__ normals,
                      unsigned char* __restrict__ valid, long long M, float radius, float tol,
                      float max_step, int max_iters) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < M;
       i += (long long)gridDim.x * blockDim.x) {
    float x = points[3 * i], y = points[3 * i + 1], z = points[3 * i + 2];
    float gx = 0.f, gy = 0.f, gz = 0.f;
    bool still = true;
    for (int it = 0;; ++it) {
      const float n = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z)));
      const float f = __fsub_rn(n, radius);
      if (n > 0.f) { gx = __fdiv_rn(x, n); gy = __fdiv_rn(y, n); gz = __fdiv_rn(z, n); }
      else { gx = gy = gz = 0.f; }
      still = fabsf(f) > tol;
      if (!still || it == max_iters) break;
      const float ss = __fadd_rn(__fadd_rn(__fmul_rn(gx, gx), __fmul_rn(gy, gy)), __fmul_rn(gz, gz));
      const float den = eps_denom_f(ss, 1.0e-17f);
      const float mx = __fmul_rn(f, __fdiv_rn(gx, den));
      const float my = __fmul_rn(f, __fdiv_rn(gy, den));
      const float mz = __fmul_rn(f, __fdiv_rn(gz, den));
      const float nrm = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(mx, mx), __fmul_rn(my, my)), __fmul_rn(mz, mz)));
      const float dn = fmaxf(nrm, 1e-15f);
      const float len = fminf(nrm, max_step);
      x = __fsub_rn(x, __fmul_rn(__fdiv_rn(mx, dn), len));
      y = __fsub_rn(y, __fmul_rn(__fdiv_rn(my, dn), len));
      z = __fsub_rn(z, __fmul_rn(__fdiv_rn(mz, dn), len));
    }
    points[3 * i] = x; points[3 * i + 1] = y; points[3 * i + 2] = z;
    normals[3 * i] = gx; normals[3 * i + 1] = gy; normals[3 * i + 2] = gz;
    valid[i] = still ? 0 : 1;
  }
}

}  // namespace isob200

using namespace isob200;

extern "C" {

size_t isob200_project_step_ws_bytes(int A) {
  return align_up((size_t)(2 + div_up(A > 0 ? A : 1, PJ_TILE)) * sizeof(unsigned));
}

// One Newton iteration of _project_points (levelset_sampling.py:313-342) on the active set.
//   points, normals : (M,3) packed, updated in place at rows act_in[0..A)
//   not_converged   : (M,) uint8/bool mask, updated at the same rows (may be NULL)
//   act_in          : (A,) int32 row ids, ascending; NULL = identity (first iteration)
//   a_dev           : NULL, or a device int holding the live row count (<= A, which then is only the
//                     upper bound used to size the launch); must not alias count_out
//   sdf (A,), grad (A,3) : SDF value and gradient at points[act_in]
//   do_update       : 0 for the final evaluation (it == proj_max_iters, :329): flags and normals
//                     are refreshed but points do not move
//   act_out (>=A ints), count_out (device int): compacted still-active rows and their number
//   next_points (>=A x 3 floats, may be NULL): updated positions of the still-active rows, compacted in
//                     the same order -- points[act_out] without a gather pass (:315 of the next iteration)
int isob200_project_step(float* points, float* normals, unsigned char* not_converged,
                         const int* act_in, int A, const int* a_dev, const float* sdf, const float* grad,
                         float tol, float max_step, int do_update, int* act_out, float* next_points,
                         int* count_out, void* ws, size_t ws_bytes, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  ISO_CHECK_ARG(A >= 0, "project_step: negative A");
  ISO_CHECK_ARG(count_out, "project_step: null count_out");
  ISO_CHECK_ARG(a_dev != count_out, "project_step: a_dev must not alias count_out");
  if (A == 0) {
    ISO_CUDA(cudaMemsetAsync(count_out, 0, sizeof(int), st));
    return ISOB200_OK;
  }
  ISO_CHECK_ARG(points && normals && sdf && grad && act_out && ws, "project_step: null pointer");
  ISO_CHECK_ARG(A < (1 << 30), "project_step: A too large");
  const size_t need = isob200_project_step_ws_bytes(A);
  if (ws_bytes < need) {
    set_error("project_step: workspace too small (%zu < %zu)", ws_bytes, need);
    return ISOB200_ERR_WORKSPACE;
  }
  ISO_CUDA(cudaMemsetAsync(ws, 0, need, st));
  const int tiles = div_up(A, PJ_TILE);
  project_step_kernel<<<tiles, PJ_THREADS, 0, st>>>(points, normals, not_converged, act_in, A, a_dev, sdf, grad,
                                                   tol, max_step, do_update, act_out, next_points,
                                                   count_out, (unsigned*)ws);
  ISO_CHECK_LAUNCH("project_step_kernel");
  return ISOB200_OK;
}

// One iteration of SphereTracing.project_points (levelset_sampling.py:733-786) on the active rays; argument
// conventions as isob200_project_step.  points (M,3) ray positions updated in place, dirs (M,3) ray directions,
// eval (M,) last SDF value per ray, grad_out (M,3) last gradient per ray (may be NULL together with grad),
// active_tol = 0.1 * proj_tolerance (:760), bound = padding + radius (:774).
int isob200_trace_step(float* points, const float* dirs, float* eval, float* grad_out, const int* act_in, int A,
                       const int* a_dev, const float* sdf, const float* grad, float active_tol, float alpha,
                       float max_step, float bound, int do_update, int* act_out, float* next_points,
                       int* count_out, void* ws, size_t ws_bytes, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  ISO_CHECK_ARG(A >= 0, "trace_step: negative A");
  ISO_CHECK_ARG(count_out, "trace_step: null count_out");
  ISO_CHECK_ARG(a_dev != count_out, "trace_step: a_dev must not alias count_out");
  if (A == 0) {
    ISO_CUDA(cudaMemsetAsync(count_out, 0, sizeof(int), st));
    return ISOB200_OK;
  }
  ISO_CHECK_ARG(points && dirs && eval && sdf && act_out && ws, "trace_step: null pointer");
  ISO_CHECK_ARG(!grad_out == !grad, "trace_step: grad and grad_out go together");
  ISO_CHECK_ARG(A < (1 << 30), "trace_step: A too large");
  const size_t need = isob200_project_step_ws_bytes(A);
  if (ws_bytes < need) {
    set_error("trace_step: workspace too small (%zu < %zu)", ws_bytes, need);
    return ISOB200_ERR_WORKSPACE;
  }
  ISO_CUDA(cudaMemsetAsync(ws, 0, need, st));
  trace_step_kernel<<<div_up(A, PJ_TILE), PJ_THREADS, 0, st>>>(points, dirs, eval, grad_out, act_in, A, a_dev, sdf,
                                                              grad, active_tol, alpha, max_step, bound, do_update,
                                                              act_out, next_points, count_out, (unsigned*)ws);
  ISO_CHECK_LAUNCH("trace_step_kernel");
  return ISOB200_OK;
}

// out_points / out_normals (>= M x 3) <- the rows with valid != 0, in order; *count_out = their number
int isob200_compact_valid(const float* points, const float* normals, const unsigned char* valid,
                          int M, float* out_points, float* out_normals, int* count_out, void* ws,
                          size_t ws_bytes, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  ISO_CHECK_ARG(M >= 0 && count_out, "compact_valid: bad arguments");
  if (M == 0) {
    ISO_CUDA(cudaMemsetAsync(count_out, 0, sizeof(int), st));
    return ISOB200_OK;
  }
  ISO_CHECK_ARG(points && valid && out_points && ws, "compact_valid: null pointer");
  ISO_CHECK_ARG(!normals == !out_normals, "compact_valid: normals and out_normals go together");
  const size_t need = isob200_project_step_ws_bytes(M);
  if (ws_bytes < need) {
    set_error("compact_valid: workspace too small (%zu < %zu)", ws_bytes, need);
    return ISOB200_ERR_WORKSPACE;
  }
  ISO_CUDA(cudaMemsetAsync(ws, 0, need, st));
  compact_valid_kernel<<<div_up(M, PJ_TILE), PJ_THREADS, 0, st>>>(points, normals, valid, M, out_points,
                                                                 out_normals, count_out, (unsigned*)ws);
  ISO_CHECK_LAUNCH("compact_valid_kernel");
  return ISOB200_OK;
}

// dst (A,3) = src[idx] ; the compacted SDF-callback input (levelset_sampling.py:315)
int isob200_gather_rows3(const float* src, const int* idx, int A, float* dst, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  if (A <= 0) return ISOB200_OK;
  ISO_CHECK_ARG(src && idx && dst, "gather_rows3: null pointer");
  gather_rows3_kernel<<<grid_for(3ll * A, 256, 8), 256, 0, st>>>(src, idx, A, dst);
  ISO_CHECK_LAUNCH("gather_rows3_kernel");
  return ISOB200_OK;
}

// Whole projection loop fused for the analytic sphere SDF f(x) = |x| - radius (BASELINE config 1).
// valid (M,) uint8 = converged mask; normals = last gradient.
int isob200_project_sphere(float* points, float* normals, unsigned char* valid, long long M,
                           float radius, float tol, float max_step, int max_iters, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  if (M <= 0) return ISOB200_OK;
  ISO_CHECK_ARG(points && normals && valid, "project_sphere: null pointer");
  ISO_CHECK_ARG(max_iters >= 0, "project_sphere: negative max_iters");
  project_sphere_kernel<<<grid_for(M, 256, 8), 256, 0, st>>>(points, normals, valid, M, radius, tol,
                                                           max_step, max_iters);
  ISO_CHECK_LAUNCH("project_sphere_kernel");
  return ISOB200_OK;
}

}  // extern "C"
