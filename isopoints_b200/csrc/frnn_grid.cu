// FRNN uniform-grid construction for sm_100a.
//
// Replaces, on the reference side:
//   external/FRNN/frnn/frnn.py:55-71          host loop computing grid params (N x .item() syncs)
//   external/FRNN/frnn/csrc/grid/grid.cu:59-184   InsertPoints{2,3}DKernel  (atomic in-cell ranks)
//   external/FRNN/external/prefix_sum             exclusive scan of cell counts (see scan.cu)
//   external/FRNN/frnn/csrc/grid/counting_sort.cu:5-125  CountingSort{2,3}DKernel
//
// Two families of entry points:
//  (1) reference-compatible primitives (insert_points / counting_sort) with the reference's
//      in-place, caller-allocated contract -- DSS's splat backward reaches into frnn._C for
//      exactly these (DSS/core/rasterizer.py:909-929);
//  (2) a fused *deterministic* build (isob200_frnn_build): cell ids -> histogram -> scan ->
//      stable LSD radix sort of (cell, point) pairs -> gather.  In-cell order is ascending
//      original index, so the grid is reproducible run to run (the reference's is not:
//      grid.cu:131 ranks points with atomicAdd).
// All kernels are HBM/L2-bound integer+gather work: coalesced loads, grid sized to the SM count.
#include "common.cuh"
#include "scan.cuh"
#include <float.h>

namespace isob200 {

// ----------------------------------------------------------------------------------------
// bounding boxes + grid parameters (device side, no host round trip)
// ----------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned f2ord(float f) {
  unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

// bbox layout: [n][0..D) = ordered-uint min, [n][4..4+D) = ordered-uint max  (8 uints per cloud)
__global__ void bbox_init_kernel(unsigned* bbox, int N) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N * 8) bbox[i] = ((i & 7) < 4) ? 0xffffffffu : 0u;
}

template <int D>
__global__ void __launch_bounds__(256)
bbox_kernel(const float* __restrict__ points, const int64_t* __restrict__ lengths, int P,
            unsigned* __restrict__ bbox) {
  const int n = blockIdx.y;
  const int len = lengths ? (int)min((long long)lengths[n], (long long)P) : P;
  const float* src = points + (size_t)n * P * D;
  float mn[D], mx[D];
#pragma unroll
  for (int d = 0; d < D; ++d) { mn[d] = FLT_MAX; mx[d] = -FLT_MAX; }
  // flat coalesced sweep over the D*len floats of this cloud; component = flat index mod D
  const long long total = (long long)len * D;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; i0 < total; i0 += 4 * stride) {
    float v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {          // four loads in flight per thread
      const long long i = i0 + u * stride;
      v[u] = i < total ? src[i] : src[i0];
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long long i = i0 + u * stride;
      const int d = (int)((i < total ? i : i0) % D);
#pragma unroll
      for (int e = 0; e < D; ++e)
        if (e == d) { mn[e] = fminf(mn[e], v[u]); mx[e] = fmaxf(mx[e], v[u]); }
    }
  }
#pragma unroll
  for (int d = 0; d < D; ++d) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      mn[d] = fminf(mn[d], __shfl_xor_sync(0xffffffffu, mn[d], o));
      mx[d] = fmaxf(mx[d], __shfl_xor_sync(0xffffffffu, mx[d], o));
    }
  }
  // one pair of atomics per block and component (a warp-level hand-off made the 2*D words a serial hot spot:
  // 20 us at 120 k points)
  __shared__ float s_mn[8][D], s_mx[8][D];
  const int w = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int d = 0; d < D; ++d) { s_mn[w][d] = mn[d]; s_mx[w][d] = mx[d]; }
  }
  __syncthreads();
  if (threadIdx.x < D && len > 0) {
    float a = s_mn[0][threadIdx.x], b = s_mx[0][threadIdx.x];
#pragma unroll
    for (int k = 1; k < 8; ++k) { a = fminf(a, s_mn[k][threadIdx.x]); b = fmaxf(b, s_mx[k][threadIdx.x]); }
    atomicMin(&bbox[n * 8 + threadIdx.x], f2ord(a));
    atomicMax(&bbox[n * 8 + 4 + threadIdx.x], f2ord(b));
  }
}

// bbox words -> floats: out[n] = (min[0..D), max[0..D)); zeros for an empty cloud
__global__ void bbox_decode_kernel(const unsigned* __restrict__ bbox, int N, int D, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * 2 * D) return;
  const int n = i / (2 * D), j = i % (2 * D);
  const bool empty = bbox[n * 8] == 0xffffffffu && bbox[n * 8 + 4] == 0u;
  out[i] = empty ? 0.f : ord2f(bbox[n * 8 + (j < D ? j : 4 + j - D)]);
}

// One thread per cloud; restates frnn.py:55-71 with the arithmetic torch performs for it on a
// CUDA device: `tensor / python_float` is evaluated as tensor * (1.0f / (float)scalar)
// (ATen div_true_kernel_cuda CPU-scalar fast path), `1 / python_float` in double then rounded
// to the fp32 params tensor, comparisons against the scalar in fp32.
template <int D>
__global__ void grid_params_kernel(const unsigned* __restrict__ bbox, const float* __restrict__ rs,
                                   int N, double radius_cell_ratio, int g_cap, float* __restrict__ params,
                                   int* __restrict__ g_max) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  constexpr int PS = (D == 3) ? ISO_G3_SIZE : ISO_G2_SIZE;
  constexpr int MAXRES = (D == 3) ? ISO_G3_MAX_RES : ISO_G2_MAX_RES;
  float gmin[D], gsize[D];
  float min_size = FLT_MAX;
  // a cloud without live rows left the box at its initial words: a one-cell grid at the origin instead of NaNs
  const bool empty = bbox[n * 8] == 0xffffffffu && bbox[n * 8 + 4] == 0u;
#pragma unroll
  for (int d = 0; d < D; ++d) {
    gmin[d] = empty ? 0.f : ord2f(bbox[n * 8 + d]);
    float gmax = empty ? 0.f : ord2f(bbox[n * 8 + 4 + d]);
    gsize[d] = __fsub_rn(gmax, gmin[d]);
    min_size = fminf(min_size, gsize[d]);
  }
  const double cell_d = (double)rs[n] / radius_cell_ratio;
  const float thresh = __fdiv_rn(min_size, (float)MAXRES);
  float delta, res[D];
  if ((float)cell_d < thresh) {
    const float cell = thresh;  // 0-dim fp32 tensor in the reference
    delta = __fdiv_rn(1.0f, cell);
#pragma unroll
    for (int d = 0; d < D; ++d) res[d] = __fadd_rn(floorf(__fdiv_rn(gsize[d], cell)), 1.0f);
  } else {
    delta = (float)(1.0 / cell_d);
    const float inv = __fdiv_rn(1.0f, (float)cell_d);
#pragma unroll
    for (int d = 0; d < D; ++d) res[d] = __fadd_rn(floorf(__fmul_rn(gsize[d], inv)), 1.0f);
  }
  float total = res[0];
#pragma unroll
  for (int d = 1; d < D; ++d) total = __fmul_rn(total, res[d]);
  float* p = params + (size_t)n * PS;
#pragma unroll
  for (int d = 0; d < D; ++d) { p[d] = gmin[d]; p[D + 1 + d] = res[d]; }
  p[D] = delta;
  p[2 * D + 1] = total;
  // `total` is a float product of up to three resolutions: with a collapsed axis (min extent 0) the reference's
  // cell-size floor (:62-63) does not engage and a tiny radius asks for more cells than an int holds.  Report
  // INT_MAX instead of an overflowed cast: the host wrapper raises (the reference would fail in its allocation).
  const int tot = (total < 1073741824.f) ? (int)total : 0x7fffffff;
  if (g_max) atomicMax(g_max, tot);
  if (tot > g_cap) {
    // the caller sized its cell table for g_cap cells before knowing the grid (isob200_frnn_grid_params_capped) and
    // this cloud needs more: hand out a one-cell grid placed where no query box reaches it, so that the build and
    // the queries stay inside the table and finish at once; g_max tells the caller to discard them and redo
#pragma unroll
    for (int d = 0; d < D; ++d) { p[d] = FLT_MAX; p[D + 1 + d] = 1.f; }
    p[D] = 1.f;
    p[2 * D + 1] = 1.f;
  }
}

// ----------------------------------------------------------------------------------------
// cell id of a point -- bit-for-bit the reference's expression (grid.cu:121-129 / :83-89)
// ----------------------------------------------------------------------------------------
template <int D>
__device__ __forceinline__ int cell_of(const float* __restrict__ pt, const float* __restrict__ prm) {
  if (D == 3) {
    const float delta = prm[ISO_G3_DELTA];
    const int rx = (int)prm[ISO_G3_RES_X], ry = (int)prm[ISO_G3_RES_Y], rz = (int)prm[ISO_G3_RES_Z];
    int gx = __float2int_rz(__fmul_rn(__fsub_rn(pt[0], prm[ISO_G3_MIN_X]), delta));
    int gy = __float2int_rz(__fmul_rn(__fsub_rn(pt[1], prm[ISO_G3_MIN_Y]), delta));
    int gz = __float2int_rz(__fmul_rn(__fsub_rn(pt[2], prm[ISO_G3_MIN_Z]), delta));
    gx = max(min(gx, rx - 1), 0);
    gy = max(min(gy, ry - 1), 0);
    gz = max(min(gz, rz - 1), 0);
    return (gx * ry + gy) * rz + gz;
  } else {
    const float delta = prm[ISO_G2_DELTA];
    const int rx = (int)prm[ISO_G2_RES_X], ry = (int)prm[ISO_G2_RES_Y];
    int gx = __float2int_rz(__fmul_rn(__fsub_rn(pt[0], prm[ISO_G2_MIN_X]), delta));
    int gy = __float2int_rz(__fmul_rn(__fsub_rn(pt[1], prm[ISO_G2_MIN_Y]), delta));
    gx = max(min(gx, rx - 1), 0);
    gy = max(min(gy, ry - 1), 0);
    return gx * ry + gy;
  }
}

// ----------------------------------------------------------------------------------------
// (1) reference-compatible primitives
// ----------------------------------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(256)
insert_points_kernel(const float* __restrict__ points, const int64_t* __restrict__ lengths,
                     const float* __restrict__ params, int* grid_cnt, int* __restrict__ grid_cell,
                     int* __restrict__ grid_idx, int N, int P, int G) {
  constexpr int PS = (D == 3) ? ISO_G3_SIZE : ISO_G2_SIZE;
  const long long total = (long long)N * P;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(i / P), p = (int)(i % P);
    if (p >= lengths[n]) continue;
    float pt[D];
#pragma unroll
    for (int d = 0; d < D; ++d) pt[d] = points[i * D + d];
    const int c = cell_of<D>(pt, params + (size_t)n * PS);
    grid_cell[i] = c;
    grid_idx[i] = atomicAdd(&grid_cnt[(size_t)n * G + c], 1);
  }
}

template <int D>
__global__ void __launch_bounds__(256)
counting_sort_kernel(const float* __restrict__ points, const int64_t* __restrict__ lengths,
                     const int* __restrict__ grid_cell, const int* __restrict__ grid_idx,
                     const int* __restrict__ grid_off, float* __restrict__ sorted_points,
                     int* __restrict__ sorted_idxs, int N, int P, int G) {
  const long long total = (long long)N * P;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(i / P), p = (int)(i % P);
    if (p >= lengths[n]) continue;
    const int c = grid_cell[i];
    const int s = grid_off[(size_t)n * G + c] + grid_idx[i];
    const size_t o = (size_t)n * P + s;
#pragma unroll
    for (int d = 0; d < D; ++d) sorted_points[o * D + d] = points[i * D + d];
    sorted_idxs[o] = p;
  }
}

// ----------------------------------------------------------------------------------------
// (2) deterministic fused build
// ----------------------------------------------------------------------------------------
// key = n*G + cell for live points, sentinel (N*G) for padded slots; also histograms cells and
// initialises the padded tail of the outputs (zeros / -1, as the reference's torch.zeros_like /
// torch.full(-1) do: frnn.py:90-91).
template <int D>
__global__ void __launch_bounds__(256)
cell_key_kernel(const float* __restrict__ points, const int64_t* __restrict__ lengths,
                const float* __restrict__ params, int N, int P, int G, unsigned* __restrict__ keys,
                int* __restrict__ cell_cnt, float* __restrict__ sorted_points,
                int* __restrict__ sorted_idxs) {
  constexpr int PS = (D == 3) ? ISO_G3_SIZE : ISO_G2_SIZE;
  const long long total = (long long)N * P;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(i / P), p = (int)(i % P);
    const long long len = lengths ? (long long)lengths[n] : (long long)P;
    if (p >= len) {
      keys[i] = (unsigned)N * (unsigned)G;
#pragma unroll
      for (int d = 0; d < D; ++d) sorted_points[i * D + d] = 0.f;
      sorted_idxs[i] = -1;
      continue;
    }
    float pt[D];
#pragma unroll
    for (int d = 0; d < D; ++d) pt[d] = points[i * D + d];
    const int c = cell_of<D>(pt, params + (size_t)n * PS);
    keys[i] = (unsigned)n * (unsigned)G + (unsigned)c;
    atomicAdd(&cell_cnt[(size_t)n * G + c], 1);  // integer sum: order-independent
  }
}

__global__ void cloud_start_kernel(const int64_t* __restrict__ lengths, int N, int P,
                                   int* __restrict__ cloud_start) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    int acc = 0;
    for (int n = 0; n < N; ++n) {
      cloud_start[n] = acc;
      long long len = lengths ? (long long)lengths[n] : (long long)P;
      acc += (int)max(0ll, min(len, (long long)P));
    }
    cloud_start[N] = acc;
  }
}

// ---- stable LSD radix sort, 8 bits per pass, (key, flat point index) pairs ----
constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_ROUNDS = 8;
constexpr int RS_TILE = RS_THREADS * RS_ROUNDS;  // 2048 elements per CTA
constexpr int RS_MAX_PASSES = 4;

// per-tile digit histogram of ONE pass, taken over the key order that pass will scatter (the
// per-tile counts depend on the order the previous pass left).  hist layout: [digit][tile]
__global__ void __launch_bounds__(RS_THREADS)
radix_hist_kernel(const unsigned* __restrict__ keys, int n, int shift, int tiles,
                  int* __restrict__ hist) {
  __shared__ int sh[256];
  sh[threadIdx.x] = 0;
  __syncthreads();
  const int base = blockIdx.x * RS_TILE;
#pragma unroll
  for (int j = 0; j < RS_ROUNDS; ++j) {
    int e = base + j * RS_THREADS + threadIdx.x;
    if (e < n) atomicAdd(&sh[(keys[e] >> shift) & 255u], 1);
  }
  __syncthreads();
  hist[(size_t)threadIdx.x * tiles + blockIdx.x] = sh[threadIdx.x];
}

// one pass: stable scatter by digit `shift/8`.  Element order inside the CTA's tile is
// warp-major, round-major, lane-minor == ascending element index, so ranks are stable.
__global__ void __launch_bounds__(RS_THREADS)
radix_scatter_kernel(const unsigned* __restrict__ keys_in, const int* __restrict__ vals_in,
                     unsigned* __restrict__ keys_out, int* __restrict__ vals_out, int n, int shift,
                     int tiles, const int* __restrict__ digit_off /* [256][tiles] scanned */) {
  __shared__ int warp_cnt[RS_WARPS][256];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < RS_WARPS * 256; i += RS_THREADS) (&warp_cnt[0][0])[i] = 0;
  __syncthreads();
  const int wbase = blockIdx.x * RS_TILE + w * (32 * RS_ROUNDS);
  unsigned key[RS_ROUNDS];
  int rank[RS_ROUNDS];
  const unsigned lt = (1u << lane) - 1u;
#pragma unroll
  for (int j = 0; j < RS_ROUNDS; ++j) {
    const int e = wbase + j * 32 + lane;
    const bool valid = e < n;
    const unsigned vmask = __ballot_sync(0xffffffffu, valid);
    rank[j] = 0;
    key[j] = 0;
    if (valid) {
      key[j] = keys_in[e];
      const unsigned dig = (key[j] >> shift) & 255u;
      const unsigned peers = __match_any_sync(vmask, dig);
      const int before = warp_cnt[w][dig];
      rank[j] = before + __popc(peers & lt);
      __syncwarp(vmask);
      if ((peers & lt) == 0) warp_cnt[w][dig] = before + __popc(peers);
      __syncwarp(vmask);
    }
  }
  __syncthreads();
  {  // per digit: exclusive scan over the warps of this CTA, seeded with the global offset
    const int dig = threadIdx.x;
    int run = digit_off[(size_t)dig * tiles + blockIdx.x];
#pragma unroll
    for (int i = 0; i < RS_WARPS; ++i) {
      int t = warp_cnt[i][dig];
      warp_cnt[i][dig] = run;
      run += t;
    }
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < RS_ROUNDS; ++j) {
    const int e = wbase + j * 32 + lane;
    if (e < n) {
      const unsigned dig = (key[j] >> shift) & 255u;
      const int pos = warp_cnt[w][dig] + rank[j];
      if (keys_out) keys_out[pos] = key[j];
      vals_out[pos] = vals_in ? vals_in[e] : e;
    }
  }
}

template <int D>
__global__ void __launch_bounds__(256)
gather_sorted_kernel(const float* __restrict__ points, const int* __restrict__ order,
                     const int* __restrict__ cloud_start, int N, int P,
                     float* __restrict__ sorted_points, int* __restrict__ sorted_idxs) {
  const int live = cloud_start[N];
  for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < live; s += gridDim.x * blockDim.x) {
    const int v = order[s];
    const int n = v / P, p = v - n * P;
    const size_t o = (size_t)n * P + (s - cloud_start[n]);
#pragma unroll
    for (int d = 0; d < D; ++d) sorted_points[o * D + d] = points[(size_t)v * D + d];
    sorted_idxs[o] = p;
  }
}

static int key_passes(long long max_key) {
  int bits = 1;
  while ((max_key >> bits) != 0) ++bits;
  return (bits + 7) / 8;
}

struct BuildWs {
  unsigned* keys_a;
  unsigned* keys_b;
  int* vals_a;
  int* vals_b;
  int* hist;
  int* cloud_start;
  void* scan_ws;
  size_t scan_ws_bytes;
  size_t total;
};

static BuildWs carve_build_ws(void* ws, int N, int P, int G) {
  const size_t n = (size_t)N * P;
  const int tiles = div_up(n ? n : 1, RS_TILE);
  BuildWs b;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    void* p = ws ? (char*)ws + off : nullptr;
    off += align_up(bytes);
    return p;
  };
  b.keys_a = (unsigned*)take(n * 4);
  b.keys_b = (unsigned*)take(n * 4);
  b.vals_a = (int*)take(n * 4);
  b.vals_b = (int*)take(n * 4);
  b.hist = (int*)take((size_t)256 * tiles * 4);
  b.cloud_start = (int*)take((size_t)(N + 1) * 4);
  size_t s1 = scan_ws_bytes(256 * tiles, 1);
  size_t s2 = scan_ws_bytes(G, N);
  b.scan_ws_bytes = s1 > s2 ? s1 : s2;
  b.scan_ws = take(b.scan_ws_bytes);
  b.total = off;
  return b;
}

template <int D>
static int frnn_build_impl(const float* points, const int64_t* lengths, const float* params, int N,
                           int P, int G, int* cell_off, float* sorted_points, int* sorted_idxs,
                           void* ws, size_t ws_bytes, cudaStream_t stream) {
  BuildWs b = carve_build_ws(ws, N, P, G);
  if (ws == nullptr || ws_bytes < b.total) {
    set_error("frnn_build: workspace too small (%zu < %zu)", ws_bytes, b.total);
    return ISOB200_ERR_WORKSPACE;
  }
  const long long n = (long long)N * P;
  if (n == 0) return ISOB200_OK;
  ISO_CHECK_ARG((long long)N * G + 1 < (1ll << 32), "frnn_build: N*G too large for 32-bit keys");
  ISO_CHECK_ARG(n < (1ll << 31), "frnn_build: N*P too large");
  // cell_off doubles as the histogram (scanned in place)
  ISO_CUDA(cudaMemsetAsync(cell_off, 0, (size_t)N * G * sizeof(int), stream));
  cell_key_kernel<D><<<grid_for(n, 256, 8), 256, 0, stream>>>(points, lengths, params, N, P, G,
                                                              b.keys_a, cell_off, sorted_points,
                                                              sorted_idxs);
  ISO_CHECK_LAUNCH("cell_key_kernel");
  cloud_start_kernel<<<1, 32, 0, stream>>>(lengths, N, P, b.cloud_start);
  ISO_CHECK_LAUNCH("cloud_start_kernel");
  int rc = exclusive_scan_i32(cell_off, cell_off, G, N, G, G, b.scan_ws, b.scan_ws_bytes, stream);
  if (rc) return rc;

  const int passes = key_passes((long long)N * G);
  const int tiles = div_up(n, RS_TILE);
  const unsigned* kin = b.keys_a;
  unsigned* kout = b.keys_b;
  const int* vin = nullptr;
  int* vout = b.vals_a;
  for (int d = 0; d < passes; ++d) {
    const bool last = (d == passes - 1);
    radix_hist_kernel<<<tiles, RS_THREADS, 0, stream>>>(kin, (int)n, 8 * d, tiles, b.hist);
    ISO_CHECK_LAUNCH("radix_hist_kernel");
    rc = exclusive_scan_i32(b.hist, b.hist, 256 * tiles, 1, 256ll * tiles, 256ll * tiles, b.scan_ws,
                            b.scan_ws_bytes, stream);
    if (rc) return rc;
    radix_scatter_kernel<<<tiles, RS_THREADS, 0, stream>>>(
        kin, vin, last ? nullptr : kout, vout, (int)n, 8 * d, tiles, b.hist);
    ISO_CHECK_LAUNCH("radix_scatter_kernel");
    unsigned* kt = (unsigned*)kin;
    kin = kout;
    kout = kt;
    vin = vout;
    vout = (vout == b.vals_a) ? b.vals_b : b.vals_a;
  }
  gather_sorted_kernel<D><<<grid_for(n, 256, 8), 256, 0, stream>>>(points, vin, b.cloud_start, N, P,
                                                                   sorted_points, sorted_idxs);
  ISO_CHECK_LAUNCH("gather_sorted_kernel");
  return ISOB200_OK;
}

static int launch_bbox(const float* points, const int64_t* lengths, int N, int P, int D, unsigned* bbox,
                       cudaStream_t stream) {
  int bx = grid_for((long long)P * D, 256, 4);
  bx = min(bx, 4 * kNumSMs);
  if (N > 1) bx = max(1, bx / N);
  if (D == 3) bbox_kernel<3><<<dim3(bx, N), 256, 0, stream>>>(points, lengths, P, bbox);
  else bbox_kernel<2><<<dim3(bx, N), 256, 0, stream>>>(points, lengths, P, bbox);
  ISO_CHECK_LAUNCH("bbox_kernel");
  return ISOB200_OK;
}

}  // namespace isob200

using namespace isob200;

extern "C" {

// params: (N, 8) for D=3 / (N, 6) for D=2, fp32, reference slot layout (grid.h:5-24).
// g_max (device int, may be null): receives max_n grid_total (the reference's G, frnn.py:70-71).
// ws: at least 32*N + 4 bytes.
int isob200_frnn_grid_params_capped(const float* points, const int64_t* lengths, const float* rs, int N,
                                    int P, int D, double radius_cell_ratio, int g_cap, float* params, int* g_max,
                                    void* ws, size_t ws_bytes, void* stream_);

int isob200_frnn_grid_params(const float* points, const int64_t* lengths, const float* rs, int N,
                             int P, int D, double radius_cell_ratio, float* params, int* g_max,
                             void* ws, size_t ws_bytes, void* stream_) {
  return isob200_frnn_grid_params_capped(points, lengths, rs, N, P, D, radius_cell_ratio, 0x7fffffff, params, g_max,
                                         ws, ws_bytes, stream_);
}

// The same with a cell budget: for a caller that allocates its (N, g_cap) cell table BEFORE reading g_max back (no
// host round trip between the parameters and the build).  A cloud whose grid needs more than g_cap cells gets a
// one-cell grid that no query reaches (build and queries stay in bounds, find nothing); g_max still receives the
// true size, so the caller learns afterwards that the results must be discarded and redone with the real G.
int isob200_frnn_grid_params_capped(const float* points, const int64_t* lengths, const float* rs, int N,
                                    int P, int D, double radius_cell_ratio, int g_cap, float* params, int* g_max,
                                    void* ws, size_t ws_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  ISO_CHECK_ARG(g_cap >= 1, "frnn_grid_params: the cell budget must be positive");
  ISO_CHECK_ARG(D == 2 || D == 3, "frnn_grid_params: only D=2/3 supported (got %d)", D);
  ISO_CHECK_ARG(N >= 0 && P >= 0, "frnn_grid_params: negative size");
  if (N == 0) return ISOB200_OK;
  ISO_CHECK_ARG(points && rs && params && ws, "frnn_grid_params: null pointer");
  if (ws_bytes < (size_t)N * 32) {
    set_error("frnn_grid_params: workspace too small");
    return ISOB200_ERR_WORKSPACE;
  }
  unsigned* bbox = (unsigned*)ws;
  bbox_init_kernel<<<div_up(N * 8, 256), 256, 0, stream>>>(bbox, N);
  ISO_CHECK_LAUNCH("bbox_init_kernel");
  if (g_max) ISO_CUDA(cudaMemsetAsync(g_max, 0, sizeof(int), stream));
  if (P > 0) {
    const int rc = launch_bbox(points, lengths, N, P, D, bbox, stream);
    if (rc != ISOB200_OK) return rc;
  }
  if (D == 3)
    grid_params_kernel<3><<<div_up(N, 64), 64, 0, stream>>>(bbox, rs, N, radius_cell_ratio, g_cap, params, g_max);
  else
    grid_params_kernel<2><<<div_up(N, 64), 64, 0, stream>>>(bbox, rs, N, radius_cell_ratio, g_cap, params, g_max);
  ISO_CHECK_LAUNCH("grid_params_kernel");
  return ISOB200_OK;
}

// Axis-aligned bounding box of the live rows of each cloud: out (N, 2, D) = min, max (zeros for an empty cloud);
// lengths (N) int64 on the device, may be null (= P).  What `points.min(0) / .max(0)` of levelset_sampling.py:254
// gives on the filtered cloud, without the host knowing the survivor count.  ws: at least 32*N bytes.
int isob200_points_bbox(const float* points, const int64_t* lengths, int N, int P, int D, float* out, void* ws,
                        size_t ws_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  ISO_CHECK_ARG(D == 2 || D == 3, "points_bbox: only D=2/3 supported (got %d)", D);
  ISO_CHECK_ARG(N >= 0 && P >= 0, "points_bbox: negative size");
  if (N == 0) return ISOB200_OK;
  ISO_CHECK_ARG(points && out && ws, "points_bbox: null pointer");
  if (ws_bytes < (size_t)N * 32) {
    set_error("points_bbox: workspace too small");
    return ISOB200_ERR_WORKSPACE;
  }
  unsigned* bbox = (unsigned*)ws;
  bbox_init_kernel<<<div_up(N * 8, 256), 256, 0, stream>>>(bbox, N);
  ISO_CHECK_LAUNCH("bbox_init_kernel");
  if (P > 0) {
    const int rc = launch_bbox(points, lengths, N, P, D, bbox, stream);
    if (rc != ISOB200_OK) return rc;
  }
  bbox_decode_kernel<<<div_up(N * 2 * D, 128), 128, 0, stream>>>(bbox, N, D, out);
  ISO_CHECK_LAUNCH("bbox_decode_kernel");
  return ISOB200_OK;
}

// == frnn._C.insert_points_cuda (grid.cu:135-184): in place on caller tensors.
int isob200_frnn_insert_points(const float* points, const int64_t* lengths, const float* params,
                               int* grid_cnt, int* grid_cell, int* grid_idx, int N, int P, int D,
                               int G, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  ISO_CHECK_ARG(D == 2 || D == 3, "for now only 2D and 3D are supported");
  if ((long long)N * P == 0) return ISOB200_OK;
  ISO_CHECK_ARG(points && lengths && params && grid_cnt && grid_cell && grid_idx,
                "insert_points: null pointer");
  const int g = grid_for((long long)N * P, 256, 8);
  if (D == 3)
    insert_points_kernel<3><<<g, 256, 0, stream>>>(points, lengths, params, grid_cnt, grid_cell, grid_idx, N, P, G);
  else
    insert_points_kernel<2><<<g, 256, 0, stream>>>(points, lengths, params, grid_cnt, grid_cell, grid_idx, N, P, G);
  ISO_CHECK_LAUNCH("insert_points_kernel");
  return ISOB200_OK;
}

// == frnn._C.counting_sort_cuda (counting_sort.cu:73-125)
int isob200_frnn_counting_sort(const float* points, const int64_t* lengths, const int* grid_cell,
                               const int* grid_idx, const int* grid_off, float* sorted_points,
                               int* sorted_idxs, int N, int P, int D, int G, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  ISO_CHECK_ARG(D == 2 || D == 3, "for now only 2D and 3D are supported");
  if ((long long)N * P == 0) return ISOB200_OK;
  ISO_CHECK_ARG(points && lengths && grid_cell && grid_idx && grid_off && sorted_points && sorted_idxs,
                "counting_sort: null pointer");
  const int g = grid_for((long long)N * P, 256, 8);
  if (D == 3)
    counting_sort_kernel<3><<<g, 256, 0, stream>>>(points, lengths, grid_cell, grid_idx, grid_off, sorted_points, sorted_idxs, N, P, G);
  else
    counting_sort_kernel<2><<<g, 256, 0, stream>>>(points, lengths, grid_cell, grid_idx, grid_off, sorted_points, sorted_idxs, N, P, G);
  ISO_CHECK_LAUNCH("counting_sort_kernel");
  return ISOB200_OK;
}

size_t isob200_frnn_build_ws_bytes(int N, int P, int G) {
  return carve_build_ws(nullptr, N, P, G).total;
}

// Fused deterministic grid build.  Outputs (caller allocated, fully written):
//   cell_off (N,G) int32 exclusive offsets, sorted_points (N,P,D), sorted_idxs (N,P) int32.
int isob200_frnn_build(const float* points, const int64_t* lengths, const float* params, int N, int P,
                       int D, int G, int* cell_off, float* sorted_points, int* sorted_idxs, void* ws,
                       size_t ws_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  ISO_CHECK_ARG(D == 2 || D == 3, "for now only 2D and 3D are supported");
  ISO_CHECK_ARG(N >= 0 && P >= 0 && G > 0, "frnn_build: bad sizes N=%d P=%d G=%d", N, P, G);
  if (N == 0) return ISOB200_OK;
  ISO_CHECK_ARG(points || P == 0, "frnn_build: null points");
  ISO_CHECK_ARG(params && cell_off && (P == 0 || (sorted_points && sorted_idxs)), "frnn_build: null pointer");
  if (P == 0) {
    ISO_CUDA(cudaMemsetAsync(cell_off, 0, (size_t)N * G * sizeof(int), stream));
    return ISOB200_OK;
  }
  if (D == 3)
    return frnn_build_impl<3>(points, lengths, params, N, P, G, cell_off, sorted_points, sorted_idxs, ws, ws_bytes, stream);
  return frnn_build_impl<2>(points, lengths, params, N, P, G, cell_off, sorted_points, sorted_idxs, ws, ws_bytes, stream);
}

}  // extern "C"
