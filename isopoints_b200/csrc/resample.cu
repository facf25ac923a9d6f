// Uniform resampling step (weighted tangential repulsion) for sm_100a.
//
// Replaces one sample_iter of UniformProjection.resample
// (DSS/models/levelset_sampling.py:268-284): two frnn_gather materialisations of (N,P,K,3)
// tensors plus ~20 elementwise/reduction kernels, by ONE kernel:
//   for each point p with neighbours j = idx[p, k] (k-th nearest, self column already dropped)
//     diff_k  = p - p_j                         (idx < 0: p_j = 0, weight forced to 0)
//     w_k     = exp(-|diff_k|^2 * inv_sigma)    inv_sigma = num_points / bbox_diagonal
//     proj_k  = diff_k - (diff_k . n_j) n_j     (tangent plane of the NEIGHBOUR's unit normal)
//     move    = (sum_k w_k + 1) * sum_k w_k proj_k / eps_denom(sum_k w_k)
//     out     = p + move
// A group of 8 lanes serves one point (lane = neighbour slot, K > 8 loops): the 8 random
// 24-byte neighbour reads (xyz + normal) are issued in parallel and hit L2 (the cloud is
// 2.4 MB at 200 k points); sums are reduced with xor-shuffles.  Compulsory HBM traffic per
// point: xyz 12 + K*4 idx + 12 out (normals/neighbour reads are L2 traffic).
#include "common.cuh"
#include <math.h>

namespace isob200 {

__device__ __forceinline__ float eps_denom_r(float x, float eps) {
  const float s = (x > 0.f) ? 1.f : ((x < 0.f) ? -1.f : 1.f);
  return s * fmaxf(fabsf(x), eps);
}

template <typename IdxT>
__global__ void __launch_bounds__(256)
resample_step_kernel(const float* __restrict__ q_points, const float* __restrict__ points,
                     const float* __restrict__ normals, const IdxT* __restrict__ idxs, int idx_stride,
                     int k_offset, const float* __restrict__ inv_sigma, int N, int Pq, int P, int K,
                     float* __restrict__ out) {
  constexpr int GW = 8;
  const int lane = threadIdx.x & 31;
  const int gl = lane & (GW - 1);
  const long long total = (long long)N * Pq;
  const long long ngroups = (long long)gridDim.x * (256 / GW);
  for (long long item = (long long)blockIdx.x * (256 / GW) + threadIdx.x / GW; item < total;
       item += ngroups) {
    const int n = (int)(item / Pq);
    const float isg = inv_sigma[n];
    const float* pts = points + (size_t)n * P * 3;
    const float* nrm = normals + (size_t)n * P * 3;
    const float px = q_points[item * 3 + 0], py = q_points[item * 3 + 1], pz = q_points[item * 3 + 2];
    float sw = 0.f, sx = 0.f, sy = 0.f, sz = 0.f;
    for (int k = gl; k < K; k += GW) {
      const long long j = (long long)idxs[item * idx_stride + k_offset + k];
      if (j >= 0) {
        const float dx = px - pts[3 * j + 0], dy = py - pts[3 * j + 1], dz = pz - pts[3 * j + 2];
        const float nx = nrm[3 * j + 0], ny = nrm[3 * j + 1], nz = nrm[3 * j + 2];
        const float d2 = dx * dx + dy * dy + dz * dz;
        const float w = expf(-d2 * isg);
        const float dn = dx * nx + dy * ny + dz * nz;
        sw += w;
        sx += w * (dx - dn * nx);
        sy += w * (dy - dn * ny);
        sz += w * (dz - dn * nz);
      }
    }
#pragma unroll
    for (int o = GW / 2; o > 0; o >>= 1) {
      sw += __shfl_xor_sync(0xffffffffu, sw, o);
      sx += __shfl_xor_sync(0xffffffffu, sx, o);
      sy += __shfl_xor_sync(0xffffffffu, sy, o);
      sz += __shfl_xor_sync(0xffffffffu, sz, o);
    }
    if (gl < 3) {
      const float den = eps_denom_r(sw, 1e-17f);
      const float dens = sw + 1.0f;
      const float s = (gl == 0) ? sx : (gl == 1 ? sy : sz);
      const float p = (gl == 0) ? px : (gl == 1 ? py : pz);
      out[item * 3 + gl] = p + dens * s / den;
    }
  }
}

// F.normalize(x, dim=-1) with eps = 1e-12 on (M,3) rows  (levelset_sampling.py:259)
__global__ void __launch_bounds__(256)
normalize_rows3_kernel(const float* __restrict__ x, long long M, float eps, float* __restrict__ out) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < M;
       i += (long long)gridDim.x * blockDim.x) {
    const float a = x[3 * i], b = x[3 * i + 1], c = x[3 * i + 2];
    const float n = fmaxf(sqrtf(a * a + b * b + c * c), eps);
    out[3 * i] = a / n; out[3 * i + 1] = b / n; out[3 * i + 2] = c / n;
  }
}

}  // namespace isob200

using namespace isob200;

extern "C" {

// One resampling move (levelset_sampling.py:268-284).
//   q_points        : (N,Pq,3) the points to move (== points for the reference's self-query; a
//                     rank's shard of the cloud in the point-sharded multi-GPU path)
//   points, normals : (N,P,3) the neighbour cloud; normals must already be unit length
//   idxs            : (N,Pq,idx_stride) neighbour ids into `points` (-1 = none), int64 or int32;
//                     the K columns starting at k_offset are used (k_offset = 1 drops the
//                     "self" column the way `idxs[..., 1:]` does, :136)
//   inv_sigma       : (N,) device floats  num_points / diag  (:256)
//   out             : (N,Pq,3) moved points (may not alias points / q_points)
int isob200_resample_step(const float* q_points, const float* points, const float* normals,
                          const void* idxs, int idx_is_i64, int idx_stride, int k_offset,
                          const float* inv_sigma, int N, int Pq, int P, int K, float* out,
                          void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  ISO_CHECK_ARG(N >= 0 && P >= 0 && Pq >= 0 && K >= 0 && k_offset >= 0 && k_offset + K <= idx_stride,
                "resample_step: bad sizes");
  if ((long long)N * Pq == 0) return ISOB200_OK;
  ISO_CHECK_ARG(q_points && points && normals && idxs && inv_sigma && out, "resample_step: null pointer");
  ISO_CHECK_ARG(points != out && q_points != out, "resample_step: out must not alias the inputs");
  const long long groups = (long long)N * Pq;
  long long need = (groups + 31) / 32;
  const long long cap = (long long)kNumSMs * 8 * 4;
  const int blocks = (int)(need < cap ? need : cap);
  if (idx_is_i64)
    resample_step_kernel<int64_t><<<blocks, 256, 0, st>>>(q_points, points, normals, (const int64_t*)idxs,
                                                         idx_stride, k_offset, inv_sigma, N, Pq, P, K, out);
  else
    resample_step_kernel<int><<<blocks, 256, 0, st>>>(q_points, points, normals, (const int*)idxs, idx_stride,
                                                     k_offset, inv_sigma, N, Pq, P, K, out);
  ISO_CHECK_LAUNCH("resample_step_kernel");
  return ISOB200_OK;
}

int isob200_normalize_rows3(const float* x, long long M, float eps, float* out, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  if (M <= 0) return ISOB200_OK;
  ISO_CHECK_ARG(x && out, "normalize_rows3: null pointer");
  normalize_rows3_kernel<<<grid_for(M, 256, 8), 256, 0, st>>>(x, M, eps, out);
  ISO_CHECK_LAUNCH("normalize_rows3_kernel");
  return ISOB200_OK;
}

}  // extern "C"
