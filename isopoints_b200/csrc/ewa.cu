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

// MUFU approximations (<= 2 ulp): the whole chain stays ~1e-6 relative, the parity bar is 1e-4
__device__ __forceinline__ float rcp_fast(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float sqrt_fast(float x) { float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float rsqrt_fast(float x) { float r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }

struct PointParams { float rx, ry, ea, eb, ec, sk; };

// One point of _get_per_point_info.  M: the view's 4x4 full projection in shared memory (p_hom @ M).
__device__ __forceinline__ PointParams point_params(float x, float y, float z, float nx, float ny, float nz,
                                                    float hk, const float* M, float pixel_var, float cutoff) {
  // ---- WJk (rasterizer.py:456-485) ----
  const float xv = x * M[0] + y * M[4] + z * M[8] + M[12];
  const float yv = x * M[1] + y * M[5] + z * M[9] + M[13];
  const float t = x * M[3] + y * M[7] + z * M[11] + M[15];
  const float it = rcp_fast(eps_denom(t));
  const float nit2 = -rcp_fast(eps_denom(t * t));
  const float j30 = nit2 * xv, j31 = nit2 * yv;
  float w0[3], w1[3];   // columns of WJk = M[:3, :] @ Jk
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    w0[r] = M[4 * r] * it + M[4 * r + 3] * j30;
    w1[r] = M[4 * r + 1] * it + M[4 * r + 3] * j31;
  }
  // ---- tangent frame (:393-398): unit u0, u1 perpendicular to the normal, u1 = n x u0 ----
  const float ax = fabsf(nx), ay = fabsf(ny), az = fabsf(nz);
  float u0x, u0y, u0z;   // n x e, e = the axis the normal is least aligned with
  if (ax <= ay && ax <= az) { u0x = 0.f; u0y = nz; u0z = -ny; }
  else if (ay <= az)        { u0x = -nz; u0y = 0.f; u0z = nx; }
  else                      { u0x = ny; u0y = -nx; u0z = 0.f; }
  // F.normalize: v / max(|v|, 1e-12) = v * rsqrt(max(|v|^2, 1e-24))
  const float inv0 = rsqrt_fast(fmaxf(u0x * u0x + u0y * u0y + u0z * u0z, 1e-24f));
  u0x *= inv0; u0y *= inv0; u0z *= inv0;
  // |n x u0| = |n| for a unit u0 perpendicular to n: one rsqrt normalises both u1 and the normal
  const float inv_n = rsqrt_fast(fmaxf(nx * nx + ny * ny + nz * nz, 1e-24f));
  const float u1x = (ny * u0z - nz * u0y) * inv_n, u1y = (nz * u0x - nx * u0z) * inv_n,
              u1z = (nx * u0y - ny * u0x) * inv_n;
  // ---- Mk = Sk WJk (2x2), Vk = WJk^T (h Sk^T Sk) WJk = h Mk^T Mk (:423-424) ----
  const float m00 = u0x * w0[0] + u0y * w0[1] + u0z * w0[2];
  const float m01 = u0x * w1[0] + u0y * w1[1] + u0z * w1[2];
  const float m10 = u1x * w0[0] + u1y * w0[1] + u1z * w0[2];
  const float m11 = u1x * w1[0] + u1y * w1[1] + u1z * w1[2];
  // det Mk = (u0 x u1) . (w0 x w1) = n_hat . (w0 x w1): less cancellation than m00 m11 - m01 m10 at
  // grazing angles
  const float det_mk = (nx * (w0[1] * w1[2] - w0[2] * w1[1]) + ny * (w0[2] * w1[0] - w0[0] * w1[2]) +
                        nz * (w0[0] * w1[1] - w0[1] * w1[0])) * inv_n;
  const float v00 = hk * (m00 * m00 + m10 * m10);
  const float v01 = hk * (m00 * m01 + m10 * m11);
  const float v11 = hk * (m01 * m01 + m11 * m11);
  // ---- variance = Vk + sigma px^2 I (:429-432); det and inverse (:528-529) ----
  const float a = v00 + pixel_var, d = v11 + pixel_var;
  // det(Vk + s I) = det Vk + s tr Vk + s^2 with det Vk = (h det Mk)^2: a sum of non-negative terms,
  // where a d - b^2 cancels for splats seen at a grazing angle
  const float hd = hk * det_mk;
  const float det = hd * hd + pixel_var * (v00 + v11) + pixel_var * pixel_var;
  const float idet = rcp_fast(det);
  PointParams o;
  o.ea = d * idet; o.eb = -2.f * v01 * idet; o.ec = a * idet;   // ellipse a x^2 + b xy + c y^2
  // ---- axis-aligned radii (:489-512); 4ac - b^2 = 4 det(variance^-1) = 4 / det ----
  const float iden = rcp_fast(eps_denom(4.f * idet));
  o.ry = sqrt_fast(eps_sqrt(4.f * o.ea * cutoff * iden));
  o.rx = sqrt_fast(eps_sqrt(4.f * o.ec * cutoff * iden));
  // ---- scaler = |det Mk| / (2 pi sqrt(det variance)) (:553-554) ----
  o.sk = fabsf(det_mk) * rcp_fast(eps_denom(sqrt_fast(eps_sqrt(det * 39.478417604357434f))));
  return o;
}

// Four consecutive points per thread: every global access is a 16-byte vector (3 float4 = 4 xyz triples).
// Needs 16-byte aligned arrays; points [4 * (P / 4), P) and unaligned inputs take the scalar kernel.
__global__ void __launch_bounds__(256)
point_params_x4_kernel(const float4* __restrict__ points, const float4* __restrict__ normals,
                       const long long* __restrict__ first_idx, int n_views, long long groups,
                       const float* __restrict__ proj, int proj_views, const float4* __restrict__ h,
                       float pixel_var, float cutoff, float4* __restrict__ radii, float4* __restrict__ ellipse,
                       float4* __restrict__ cutoff_out, float4* __restrict__ scaler) {
  __shared__ float s_m[MAX_VIEWS * 16];
  __shared__ long long s_first[MAX_VIEWS];
  for (int v = threadIdx.x; v < proj_views * 16; v += blockDim.x) s_m[v] = proj[v];
  for (int v = threadIdx.x; v < n_views; v += blockDim.x) s_first[v] = first_idx[v];
  __syncthreads();
  const float4 cut4 = make_float4(cutoff, cutoff, cutoff, cutoff);
  for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < groups;
       g += (long long)gridDim.x * blockDim.x) {
    const float4 p0 = __ldg(points + 3 * g), p1 = __ldg(points + 3 * g + 1), p2 = __ldg(points + 3 * g + 2);
    const float4 n0 = __ldg(normals + 3 * g), n1 = __ldg(normals + 3 * g + 1), n2 = __ldg(normals + 3 * g + 2);
    const float4 h4 = __ldg(h + g);
    const float px[4] = {p0.x, p0.w, p1.z, p2.y}, py[4] = {p0.y, p1.x, p1.w, p2.z}, pz[4] = {p0.z, p1.y, p2.x, p2.w};
    const float qx[4] = {n0.x, n0.w, n1.z, n2.y}, qy[4] = {n0.y, n1.x, n1.w, n2.z}, qz[4] = {n0.z, n1.y, n2.x, n2.w};
    const float hh[4] = {h4.x, h4.y, h4.z, h4.w};
    const int b0 = view_of(s_first, n_views, 4 * g), b3 = view_of(s_first, n_views, 4 * g + 3);
    PointParams o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int b = (b0 == b3) ? b0 : view_of(s_first, n_views, 4 * g + j);
      o[j] = point_params(px[j], py[j], pz[j], qx[j], qy[j], qz[j], hh[j], s_m + (proj_views > 1 ? b : 0) * 16,
                          pixel_var, cutoff);
    }
    radii[2 * g] = make_float4(o[0].rx, o[0].ry, o[1].rx, o[1].ry);
    radii[2 * g + 1] = make_float4(o[2].rx, o[2].ry, o[3].rx, o[3].ry);
    ellipse[3 * g] = make_float4(o[0].ea, o[0].eb, o[0].ec, o[1].ea);
    ellipse[3 * g + 1] = make_float4(o[1].eb, o[1].ec, o[2].ea, o[2].eb);
    ellipse[3 * g + 2] = make_float4(o[2].ec, o[3].ea, o[3].eb, o[3].ec);
    cutoff_out[g] = cut4;
    scaler[g] = make_float4(o[0].sk, o[1].sk, o[2].sk, o[3].sk);
  }
}

// scalar path: points [begin, P)
__global__ void __launch_bounds__(256)
point_params_kernel(const float* __restrict__ points, const float* __restrict__ normals,
                    const long long* __restrict__ first_idx, int n_views, long long begin, long long P,
                    const float* __restrict__ proj, int proj_views, const float* __restrict__ h,
                    float pixel_var, float cutoff, float* __restrict__ radii, float* __restrict__ ellipse,
                    float* __restrict__ cutoff_out, float* __restrict__ scaler) {
  __shared__ float s_m[MAX_VIEWS * 16];
  __shared__ long long s_first[MAX_VIEWS];
  for (int v = threadIdx.x; v < proj_views * 16; v += blockDim.x) s_m[v] = proj[v];
  for (int v = threadIdx.x; v < n_views; v += blockDim.x) s_first[v] = first_idx[v];
  __syncthreads();
  for (long long i = begin + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < P;
       i += (long long)gridDim.x * blockDim.x) {
    const int b = view_of(s_first, n_views, i);
    const PointParams o = point_params(points[3 * i], points[3 * i + 1], points[3 * i + 2], normals[3 * i],
                                       normals[3 * i + 1], normals[3 * i + 2], h[i],
                                       s_m + (proj_views > 1 ? b : 0) * 16, pixel_var, cutoff);
    radii[2 * i] = o.rx;
    radii[2 * i + 1] = o.ry;
    ellipse[3 * i] = o.ea;
    ellipse[3 * i + 1] = o.eb;
    ellipse[3 * i + 2] = o.ec;
    cutoff_out[i] = cutoff;
    scaler[i] = o.sk;
  }
}

// mask[i] = znear <= z_view <= zfar  [and n_view.z < 0]; kept[b] += mask (per view).
//   w2v (views,4,4): world-to-view matrix, row-vector convention; nmat (views,3,3): the matrix normals are
//   multiplied with (pytorch3d Transform3d.transform_normals: inverse(w2v)[:3,:3]^T); null = no culling.
// Four consecutive points per thread (float4 loads, one uchar4 store) when `vec`; else one point per thread.
struct MaskCams {
  float v[MAX_VIEWS * 8];    // columns 2 and 3 of w2v
  float n[MAX_VIEWS * 3];    // column 2 of nmat
  long long first[MAX_VIEWS];
  int kept[MAX_VIEWS];
};

__device__ __forceinline__ bool renderable(const MaskCams& c, int cam, bool cull, float x, float y, float z,
                                           float nx, float ny, float nz, float znear, float zfar) {
  const float* V = c.v + cam * 8;
  // exact division: the depth test is a decision, keep it as close to the reference's arithmetic as possible
  const float zv = (x * V[0] + y * V[2] + z * V[4] + V[6]) / (x * V[1] + y * V[3] + z * V[5] + V[7]);
  bool keep = (zv >= znear) && (zv <= zfar);
  if (cull) {
    const float* Nm = c.n + cam * 3;
    keep = keep && (nx * Nm[0] + ny * Nm[1] + nz * Nm[2] < 0.f);
  }
  return keep;
}

template <bool VEC>
__global__ void __launch_bounds__(256)
renderable_mask_kernel(const float* __restrict__ points, const float* __restrict__ normals,
                       const long long* __restrict__ first_idx, int n_views, long long begin, long long P,
                       const float* __restrict__ w2v, const float* __restrict__ nmat, int cam_views,
                       float znear, float zfar, unsigned char* __restrict__ mask, int* __restrict__ kept) {
  __shared__ MaskCams c;
  for (int v = threadIdx.x; v < cam_views * 8; v += blockDim.x) {
    const int cam = v >> 3, r = (v & 7) >> 1, col = 2 + (v & 1);
    c.v[v] = w2v[cam * 16 + r * 4 + col];
  }
  if (nmat)
    for (int v = threadIdx.x; v < cam_views * 3; v += blockDim.x) c.n[v] = nmat[(v / 3) * 9 + (v % 3) * 3 + 2];
  for (int v = threadIdx.x; v < n_views; v += blockDim.x) {
    c.first[v] = first_idx[v];
    c.kept[v] = 0;
  }
  __syncthreads();
  const bool cull = nmat != nullptr;
  const long long per = (long long)gridDim.x * blockDim.x;
  const long long items = VEC ? P / 4 : P - begin;   // VEC: groups of 4 from point 0; scalar: points from `begin`
  for (long long i0 = (long long)blockIdx.x * blockDim.x; i0 < items; i0 += per) {
    const long long it = i0 + threadIdx.x;
    int cnt = 0, b = 0;
    bool straddle = false;
    if (it < items) {
      if (VEC) {
        const float4* p4 = reinterpret_cast<const float4*>(points) + 3 * it;
        const float4 p0 = __ldg(p4), p1 = __ldg(p4 + 1), p2 = __ldg(p4 + 2);
        float4 n0 = make_float4(0.f, 0.f, 0.f, 0.f), n1 = n0, n2 = n0;
        if (cull) {
          const float4* n4 = reinterpret_cast<const float4*>(normals) + 3 * it;
          n0 = __ldg(n4); n1 = __ldg(n4 + 1); n2 = __ldg(n4 + 2);
        }
        const float px[4] = {p0.x, p0.w, p1.z, p2.y}, py[4] = {p0.y, p1.x, p1.w, p2.z}, pz[4] = {p0.z, p1.y, p2.x, p2.w};
        const float qx[4] = {n0.x, n0.w, n1.z, n2.y}, qy[4] = {n0.y, n1.x, n1.w, n2.z}, qz[4] = {n0.z, n1.y, n2.x, n2.w};
        b = view_of(c.first, n_views, 4 * it);
        straddle = view_of(c.first, n_views, 4 * it + 3) != b;
        unsigned char k[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int bj = straddle ? view_of(c.first, n_views, 4 * it + j) : b;
          k[j] = renderable(c, cam_views > 1 ? bj : 0, cull, px[j], py[j], pz[j], qx[j], qy[j], qz[j], znear, zfar);
          if (straddle) { if (k[j]) atomicAdd(&c.kept[bj], 1); } else cnt += k[j];
        }
        reinterpret_cast<uchar4*>(mask)[it] = make_uchar4(k[0], k[1], k[2], k[3]);
      } else {
        const long long i = begin + it;
        b = view_of(c.first, n_views, i);
        float nx = 0.f, ny = 0.f, nz = 0.f;
        if (cull) { nx = normals[3 * i]; ny = normals[3 * i + 1]; nz = normals[3 * i + 2]; }
        cnt = renderable(c, cam_views > 1 ? b : 0, cull, points[3 * i], points[3 * i + 1], points[3 * i + 2], nx, ny,
                         nz, znear, zfar);
        mask[i] = (unsigned char)cnt;
      }
    }
    // per-view survivor counts: a warp spans at most a few views; reduce over the lanes of one view
    const unsigned act = __ballot_sync(0xffffffffu, cnt > 0);
    if (cnt > 0) {
      const unsigned same = __match_any_sync(act, b);
      const int sum = __reduce_add_sync(same, cnt);
      if ((threadIdx.x & 31) == __ffs(same) - 1) atomicAdd(&c.kept[b], sum);
    }
  }
  __syncthreads();
  for (int v = threadIdx.x; v < n_views; v += blockDim.x)
    if (c.kept[v]) atomicAdd(&kept[v], c.kept[v]);
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
  cudaStream_t st = (cudaStream_t)stream;
  const uintptr_t al = (uintptr_t)points | (uintptr_t)normals | (uintptr_t)vrk_h | (uintptr_t)radii |
                       (uintptr_t)ellipse | (uintptr_t)cutoff_out | (uintptr_t)scaler;
  const long long groups = (al & 15) == 0 ? P / 4 : 0;
  if (groups > 0) {
    ewa::point_params_x4_kernel<<<grid_for(groups, 256, 4), 256, 0, st>>>(
        (const float4*)points, (const float4*)normals, (const long long*)first_idx, n_views, groups, proj, proj_views,
        (const float4*)vrk_h, pixel_var, cutoff, (float4*)radii, (float4*)ellipse, (float4*)cutoff_out,
        (float4*)scaler);
    ISO_CHECK_LAUNCH("ewa_point_params_x4_kernel");
  }
  if (4 * groups < P) {
    ewa::point_params_kernel<<<grid_for(P - 4 * groups, 256, 8), 256, 0, st>>>(
        points, normals, (const long long*)first_idx, n_views, 4 * groups, P, proj, proj_views, vrk_h, pixel_var,
        cutoff, radii, ellipse, cutoff_out, scaler);
    ISO_CHECK_LAUNCH("ewa_point_params_kernel");
  }
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
  cudaStream_t st = (cudaStream_t)stream;
  const uintptr_t al = (uintptr_t)points | (uintptr_t)normals | (uintptr_t)mask;
  const long long groups = (al & 15) == 0 ? P / 4 : 0;
  if (groups > 0) {
    ewa::renderable_mask_kernel<true><<<grid_for(groups, 256, 4), 256, 0, st>>>(
        points, normals, (const long long*)first_idx, n_views, 0, 4 * groups, w2v, nmat, cam_views, znear, zfar, mask,
        kept);
    ISO_CHECK_LAUNCH("renderable_mask_x4_kernel");
  }
  if (4 * groups < P) {
    ewa::renderable_mask_kernel<false><<<grid_for(P - 4 * groups, 256, 8), 256, 0, st>>>(
        points, normals, (const long long*)first_idx, n_views, 4 * groups, P, w2v, nmat, cam_views, znear, zfar, mask,
        kept);
    ISO_CHECK_LAUNCH("renderable_mask_kernel");
  }
  return ISOB200_OK;
}

}  // extern "C"
