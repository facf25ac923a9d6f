// Per-point EWA splat parameters and the renderable filter, for sm_100a.
//
// Replaces the PyTorch op chains that sit immediately before the splat kernel in
// DSS/core/rasterizer.py:
//   _compute_isotropic_Vrk   (:344-400)  FRNN/KNN K=7 distances -> h_k, local frame S_k
//   _compute_WJk             (:438-487)  Jacobian of the camera projection per point
//   _compute_variance_and_detMk (:402-436) V_k = WJk^T V_k^r WJk + sigma px^2 I, det(M_k)
//   _get_per_point_info      (:514-563)  det / inverse of the 2x2 variance, axis-aligned radii, scaler
//   _filter_points_with_invalid_depth (:163-218), _filter_backface_points (:124-161)
// i.e. ~30 small elementwise / batched-matmul launches with (P,3,3), (P,4,2), (P,3,2) temporaries
// (and a batched LU for det / inverse of 2x2 matrices), per view batch, 4x per training step.
// Here: one thread per point, camera matrices in shared memory, everything in registers;
// 28 B read + 28 B written per point -> HBM-bound.
//
// The reference draws the tangent frame S_k = (u0, u1) with torch.rand_like (:393-396).  Every
// quantity that leaves _get_per_point_info depends on S_k only through S_k^T S_k (the projector onto
// the tangent plane) and det(S_k WJk) (invariant under rotations of the frame, and u1 = n x u0 fixes the
// handedness), so a deterministic frame gives the same results up to rounding.
#include "common.cuh"
#include <math.h>

namespace isob200 {
namespace ewa {

constexpr int MAX_VIEWS = 64;
constexpr float EPS = 1e-17f;   // mathHelper.py:14, :20 default eps

// eps_denom (mathHelper.py:14-18): (sign(x) + [x == 0]) * max(|x|, eps)
__device__ __forceinline__ float eps_denom(float x) {
  return ((x < 0.f) ? -1.f : 1.f) * fmaxf(fabsf(x), EPS);
}
// eps_sqrt (mathHelper.py:20-25): max(|x|, eps)
__device__ __forceinline__ float eps_sqrt(float x) { return fmaxf(fabsf(x), EPS); }

// view of packed point i: last b with first_idx[b] <= i (first_idx ascending, empty views allowed)
__device__ __forceinline__ int view_of(const long long* s_first, int n_views, long long i) {
  int b = 0;
  for (int v = 1; v < n_views; ++v) b = (s_first[v] <= i) ? v : b;
  return b;
}

// h_k = clamp(0.5 * max_k sq_dist[k >= 1], 5e-5, 0.01) (rasterizer.py:374-386); dists is the padded
// (N, P1, K) output of the K = 7 self query, slot 0 being the point itself; clouds with fewer than
// K points get sq_dist = 1e-3 (:376).  Output packed (P,).
__global__ void __launch_bounds__(256)
vrk_h_kernel(const float* __restrict__ dists, const long long* __restrict__ first_idx,
             const long long* __restrict__ num_points, int n_views, long long P1, int K, long long P,
             float* __restrict__ h) {
  __shared__ long long s_first[MAX_VIEWS];
  for (int v = threadIdx.x; v < n_views; v += blockDim.x) s_first[v] = first_idx[v];
  __syncthreads();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < P;
       i += (long long)gridDim.x * blockDim.x) {
    const int b = view_of(s_first, n_views, i);
    float m;
    if (num_points[b] < K) {
      m = 1e-3f;
    } else {
      const float* d = dists + ((size_t)b * P1 + (size_t)(i - s_first[b])) * K;
      m = d[1];
      for (int k = 2; k < K; ++k) m = fmaxf(m, d[k]);
    }
    h[i] = fminf(fmaxf(0.5f * m, 5e-5f), 0.01f);
  }
}

__global__ void __launch_bounds__(256)
point_params_kernel(const float* __restrict__ points, const float* __restrict__ normals,
                    const long long* __restrict__ first_idx, int n_views, long long P,
                    const float* __restrict__ proj, int proj_views, const float* __restrict__ h,
                    float pixel_var, float cutoff, float* __restrict__ radii, float* __restrict__ ellipse,
                    float* __restrict__ cutoff_out, float* __restrict__ scaler) {
  __shared__ float s_m[MAX_VIEWS * 16];
  __shared__ long long s_first[MAX_VIEWS];
  for (int v = threadIdx.x; v < proj_views * 16; v += blockDim.x) s_m[v] = proj[v];
  for (int v = threadIdx.x; v < n_views; v += blockDim.x) s_first[v] = first_idx[v];
  __syncthreads();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < P;
       i += (long long)gridDim.x * blockDim.x) {
    const int b = view_of(s_first, n_views, i);
    const float* M = s_m + (proj_views > 1 ? b : 0) * 16;   // row-vector convention: p_hom @ M
    const float x = points[3 * i], y = points[3 * i + 1], z = points[3 * i + 2];
    // ---- WJk (rasterizer.py:456-485) ----
    const float xv = x * M[0] + y * M[4] + z * M[8] + M[12];
    const float yv = x * M[1] + y * M[5] + z * M[9] + M[13];
    const float t = x * M[3] + y * M[7] + z * M[11] + M[15];
    const float t2 = eps_denom(t * t);
    const float it = 1.f / eps_denom(t);
    const float j30 = (-1.f / t2) * xv, j31 = (-1.f / t2) * yv;
    float w0[3], w1[3];   // columns of WJk = M[:3, :] @ Jk
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      w0[r] = M[4 * r] * it + M[4 * r + 3] * j30;
      w1[r] = M[4 * r + 1] * it + M[4 * r + 3] * j31;
    }
    // ---- tangent frame (:393-398): unit u0, u1 perpendicular to the normal, u1 = n x u0 ----
    const float nx = normals[3 * i], ny = normals[3 * i + 1], nz = normals[3 * i + 2];
    const float ax = fabsf(nx), ay = fabsf(ny), az = fabsf(nz);
    float u0x, u0y, u0z;   // n x e, e = the axis the normal is least aligned with
    if (ax <= ay && ax <= az) { u0x = 0.f; u0y = nz; u0z = -ny; }
    else if (ay <= az)        { u0x = -nz; u0y = 0.f; u0z = nx; }
    else                      { u0x = ny; u0y = -nx; u0z = 0.f; }
    float inv = 1.f / fmaxf(sqrtf(u0x * u0x + u0y * u0y + u0z * u0z), 1e-12f);   // F.normalize eps
    u0x *= inv; u0y *= inv; u0z *= inv;
    float u1x = ny * u0z - nz * u0y, u1y = nz * u0x - nx * u0z, u1z = nx * u0y - ny * u0x;
    inv = 1.f / fmaxf(sqrtf(u1x * u1x + u1y * u1y + u1z * u1z), 1e-12f);
    u1x *= inv; u1y *= inv; u1z *= inv;
    // ---- Mk = Sk WJk (2x2), Vk = WJk^T (h Sk^T Sk) WJk = h Mk^T Mk (:423-424) ----
    const float m00 = u0x * w0[0] + u0y * w0[1] + u0z * w0[2];
    const float m01 = u0x * w1[0] + u0y * w1[1] + u0z * w1[2];
    const float m10 = u1x * w0[0] + u1y * w0[1] + u1z * w0[2];
    const float m11 = u1x * w1[0] + u1y * w1[1] + u1z * w1[2];
    // det Mk = (u0 x u1) . (w0 x w1) = n_hat . (w0 x w1): less cancellation than m00 m11 - m01 m10 at
    // grazing angles
    const float inv_n = 1.f / fmaxf(sqrtf(nx * nx + ny * ny + nz * nz), 1e-12f);
    const float det_mk = (nx * (w0[1] * w1[2] - w0[2] * w1[1]) + ny * (w0[2] * w1[0] - w0[0] * w1[2]) +
                          nz * (w0[0] * w1[1] - w0[1] * w1[0])) * inv_n;
    const float hk = h[i];
    const float v00 = hk * (m00 * m00 + m10 * m10);
    const float v01 = hk * (m00 * m01 + m10 * m11);
    const float v11 = hk * (m01 * m01 + m11 * m11);
    // ---- variance = Vk + sigma px^2 I (:429-432); det and inverse (:528-529) ----
    const float a = v00 + pixel_var, d = v11 + pixel_var;
    // det(Vk + s I) = det Vk + s tr Vk + s^2 with det Vk = (h det Mk)^2: a sum of non-negative terms,
    // where a d - b^2 cancels for splats seen at a grazing angle
    const float hd = hk * det_mk;
    const float det = hd * hd + pixel_var * (v00 + v11) + pixel_var * pixel_var;
    const float idet = 1.f / det;
    const float ea = d * idet, eb = -2.f * v01 * idet, ec = a * idet;   // ellipse a x^2 + b xy + c y^2
    // ---- axis-aligned radii (:489-512); 4ac - b^2 = 4 det(variance^-1) = 4 / det ----
    const float den = eps_denom(4.f * idet);
    const float ry = sqrtf(eps_sqrt(4.f * ea * cutoff / den));
    const float rx = sqrtf(eps_sqrt(4.f * ec * cutoff / den));
    // ---- scaler = |det Mk| / (2 pi sqrt(det variance)) (:553-554) ----
    const float sk = fabsf(det_mk) / eps_denom(sqrtf(eps_sqrt(det * 39.478417604357434f)));
    reinterpret_cast<float2*>(radii)[i] = make_float2(rx, ry);
    ellipse[3 * i] = ea;
    ellipse[3 * i + 1] = eb;
    ellipse[3 * i + 2] = ec;
    cutoff_out[i] = cutoff;
    scaler[i] = sk;
  }
}

// mask[i] = znear <= z_view <= zfar  [and n_view.z < 0]; kept[b] += mask (per view).
//   w2v (views,4,4): world-to-view matrix, row-vector convention; nmat (views,3,3): the matrix normals are
//   multiplied with (pytorch3d Transform3d.transform_normals: inverse(w2v)[:3,:3]^T); null = no culling.
__global__ void __launch_bounds__(256)
renderable_mask_kernel(const float* __restrict__ points, const float* __restrict__ normals,
                       const long long* __restrict__ first_idx, int n_views, long long P,
                       const float* __restrict__ w2v, const float* __restrict__ nmat, int cam_views,
                       float znear, float zfar, unsigned char* __restrict__ mask, int* __restrict__ kept) {
  __shared__ float s_v[MAX_VIEWS * 8];    // columns 2 and 3 of w2v
  __shared__ float s_n[MAX_VIEWS * 3];    // column 2 of nmat
  __shared__ long long s_first[MAX_VIEWS];
  __shared__ int s_kept[MAX_VIEWS];
  for (int v = threadIdx.x; v < cam_views * 8; v += blockDim.x) {
    const int c = v >> 3, r = (v & 7) >> 1, col = 2 + (v & 1);
    s_v[v] = w2v[c * 16 + r * 4 + col];
  }
  if (nmat)
    for (int v = threadIdx.x; v < cam_views * 3; v += blockDim.x) s_n[v] = nmat[(v / 3) * 9 + (v % 3) * 3 + 2];
  for (int v = threadIdx.x; v < n_views; v += blockDim.x) {
    s_first[v] = first_idx[v];
    s_kept[v] = 0;
  }
  __syncthreads();
  const long long per = (long long)gridDim.x * blockDim.x;
  for (long long i0 = (long long)blockIdx.x * blockDim.x; i0 < P; i0 += per) {
    const long long i = i0 + threadIdx.x;
    bool keep = false;
    int b = 0;
    if (i < P) {
      b = view_of(s_first, n_views, i);
      const int c = cam_views > 1 ? b : 0;
      const float* V = s_v + c * 8;
      const float x = points[3 * i], y = points[3 * i + 1], z = points[3 * i + 2];
      const float zv = (x * V[0] + y * V[2] + z * V[4] + V[6]) / (x * V[1] + y * V[3] + z * V[5] + V[7]);
      keep = (zv >= znear) && (zv <= zfar);
      if (nmat) {
        const float* Nm = s_n + c * 3;
        const float nzv = normals[3 * i] * Nm[0] + normals[3 * i + 1] * Nm[1] + normals[3 * i + 2] * Nm[2];
        keep = keep && (nzv < 0.f);
      }
      mask[i] = keep ? 1 : 0;
    }
    // per-view survivor counts: a warp spans at most a few views; match on the view id
    const unsigned act = __ballot_sync(0xffffffffu, keep);
    if (keep) {
      const unsigned same = __match_any_sync(act, b);
      if ((threadIdx.x & 31) == __ffs(same) - 1) atomicAdd(&s_kept[b], __popc(same));
    }
  }
  __syncthreads();
  for (int v = threadIdx.x; v < n_views; v += blockDim.x)
    if (s_kept[v]) atomicAdd(&kept[v], s_kept[v]);
}

}  // namespace ewa
}  // namespace isob200

using namespace isob200;

extern "C" {

int isob200_ewa_vrk_h(const float* sq_dists, const int64_t* first_idx, const int64_t* num_points, int n_views,
                      long long P1, int K, long long P, float* h, void* stream) {
  if (P <= 0) return ISOB200_OK;
  ISO_CHECK_ARG(sq_dists && first_idx && num_points && h, "ewa_vrk_h: null pointer");
  ISO_CHECK_ARG(n_views >= 1 && n_views <= ewa::MAX_VIEWS, "ewa_vrk_h: n_views must be in 1..%d", ewa::MAX_VIEWS);
  ISO_CHECK_ARG(K >= 2 && P1 >= 1, "ewa_vrk_h: need K >= 2 neighbour slots (slot 0 is the point itself)");
  ewa::vrk_h_kernel<<<grid_for(P, 256, 8), 256, 0, (cudaStream_t)stream>>>(
      sq_dists, (const long long*)first_idx, (const long long*)num_points, n_views, P1, K, P, h);
  ISO_CHECK_LAUNCH("ewa_vrk_h_kernel");
  return ISOB200_OK;
}

int isob200_ewa_point_params(const float* points, const float* normals, const int64_t* first_idx, int n_views,
                             long long P, const float* proj, int proj_views, const float* vrk_h,
                             float pixel_var, float cutoff, float* radii, float* ellipse, float* cutoff_out,
                             float* scaler, void* stream) {
  if (P <= 0) return ISOB200_OK;
  ISO_CHECK_ARG(points && normals && first_idx && proj && vrk_h, "ewa_point_params: null input pointer");
  ISO_CHECK_ARG(radii && ellipse && cutoff_out && scaler, "ewa_point_params: null output pointer");
  ISO_CHECK_ARG(n_views >= 1 && n_views <= ewa::MAX_VIEWS, "ewa_point_params: n_views must be in 1..%d",
                ewa::MAX_VIEWS);
  ISO_CHECK_ARG(proj_views == 1 || proj_views == n_views,
                "ewa_point_params: %d cameras for %d point clouds", proj_views, n_views);
  ISO_CHECK_ARG(((uintptr_t)radii & 7) == 0, "ewa_point_params: radii must be 8-byte aligned");
  ewa::point_params_kernel<<<grid_for(P, 256, 8), 256, 0, (cudaStream_t)stream>>>(
      points, normals, (const long long*)first_idx, n_views, P, proj, proj_views, vrk_h, pixel_var, cutoff,
      radii, ellipse, cutoff_out, scaler);
  ISO_CHECK_LAUNCH("ewa_point_params_kernel");
  return ISOB200_OK;
}

int isob200_renderable_mask(const float* points, const float* normals, const int64_t* first_idx, int n_views,
                            long long P, const float* w2v, const float* nmat, int cam_views, float znear,
                            float zfar, unsigned char* mask, int* kept, void* stream) {
  ISO_CHECK_ARG(n_views >= 1 && n_views <= ewa::MAX_VIEWS, "renderable_mask: n_views must be in 1..%d",
                ewa::MAX_VIEWS);
  ISO_CHECK_ARG(kept, "renderable_mask: null kept pointer");
  ISO_CUDA(cudaMemsetAsync(kept, 0, sizeof(int) * n_views, (cudaStream_t)stream));
  if (P <= 0) return ISOB200_OK;
  ISO_CHECK_ARG(points && first_idx && w2v && mask, "renderable_mask: null pointer");
  ISO_CHECK_ARG(!nmat || normals, "renderable_mask: back-face culling needs normals");
  ISO_CHECK_ARG(cam_views == 1 || cam_views == n_views, "renderable_mask: %d cameras for %d point clouds",
                cam_views, n_views);
  ewa::renderable_mask_kernel<<<grid_for(P, 256, 8), 256, 0, (cudaStream_t)stream>>>(
      points, normals, (const long long*)first_idx, n_views, P, w2v, nmat, cam_views, znear, zfar, mask, kept);
  ISO_CHECK_LAUNCH("renderable_mask_kernel");
  return ISOB200_OK;
}

}  // extern "C"
