// Exclusive prefix sum of int32 counters (per-cell point counts, radix digit histograms,
// per-tile splat counts). Replaces the reference's Blelloch scan
// (external/FRNN/external/prefix_sum/prefix_sum.cu:74-87, one call per cloud on the
// default stream with a cudaMalloc per call) by a batched reduce / spine / downsweep
// on the caller's stream with a caller-provided workspace. HBM-bound: 8 B per element.
#include "common.cuh"
#include "scan.cuh"

namespace isob200 {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ int warp_incl_scan(int v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += t;
  }
  return v;
}

// exclusive prefix of `v` across the block (SCAN_THREADS threads); *total gets the block sum
__device__ __forceinline__ int block_excl_scan(int v, int* total) {
  __shared__ int warp_tot[SCAN_THREADS / 32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int incl = warp_incl_scan(v, lane);
  if (lane == 31) warp_tot[w] = incl;
  __syncthreads();
  int woff = 0, tot = 0;
#pragma unroll
  for (int i = 0; i < SCAN_THREADS / 32; ++i) {
    int t = warp_tot[i];
    if (i < w) woff += t;
    tot += t;
  }
  __syncthreads();
  if (total) *total = tot;
  return woff + incl - v;
}

__global__ void __launch_bounds__(SCAN_THREADS)
scan_reduce_kernel(const int* __restrict__ in, int n, long long in_stride, int tiles,
                   int* __restrict__ tile_sums) {
  const int row = blockIdx.y, tile = blockIdx.x;
  const int* src = in + (long long)row * in_stride;
  const int base = tile * SCAN_TILE;
  int s = 0;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) {
    int j = base + i * SCAN_THREADS + threadIdx.x;  // coalesced, order irrelevant for a sum
    if (j < n) s += src[j];
  }
  int tot;
  block_excl_scan(s, &tot);
  if (threadIdx.x == 0) tile_sums[row * tiles + tile] = tot;
}

// one block per row: in-place exclusive scan of that row's tile sums
__global__ void __launch_bounds__(SCAN_THREADS)
scan_spine_kernel(int* __restrict__ tile_sums, int tiles) {
  int* row = tile_sums + (long long)blockIdx.x * tiles;
  int carry = 0;
  for (int base = 0; base < tiles; base += SCAN_THREADS) {
    int j = base + threadIdx.x;
    int v = j < tiles ? row[j] : 0;
    int tot;
    int ex = block_excl_scan(v, &tot);
    if (j < tiles) row[j] = carry + ex;
    carry += tot;
  }
}

__global__ void __launch_bounds__(SCAN_THREADS)
scan_downsweep_kernel(const int* in, int* out, int n,
                      long long in_stride, long long out_stride, int tiles,
                      const int* __restrict__ tile_offs) {
  const int row = blockIdx.y, tile = blockIdx.x;
  const int* src = in + (long long)row * in_stride;
  int* dst = out + (long long)row * out_stride;
  const int base = tile * SCAN_TILE + threadIdx.x * SCAN_ITEMS;  // blocked: keeps element order
  int v[SCAN_ITEMS];
  int s = 0;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) {
    v[i] = (base + i < n) ? src[base + i] : 0;
    s += v[i];
  }
  int ex = block_excl_scan(s, nullptr);
  if (tile_offs) ex += tile_offs[row * tiles + tile];
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) {
    if (base + i < n) dst[base + i] = ex;
    ex += v[i];
  }
}

size_t scan_ws_bytes(int n, int rows) {
  int tiles = div_up(n > 0 ? n : 1, SCAN_TILE);
  return align_up((size_t)tiles * rows * sizeof(int));
}

int exclusive_scan_i32(const int* in, int* out, int n, int rows, long long in_stride,
                       long long out_stride, void* ws, size_t ws_bytes, cudaStream_t stream) {
  if (n <= 0 || rows <= 0) return ISOB200_OK;
  const int tiles = div_up(n, SCAN_TILE);
  if (tiles == 1) {
    scan_downsweep_kernel<<<dim3(1, rows), SCAN_THREADS, 0, stream>>>(in, out, n, in_stride,
                                                                      out_stride, 1, nullptr);
    ISO_CHECK_LAUNCH("scan_downsweep");
    return ISOB200_OK;
  }
  if (ws == nullptr || ws_bytes < scan_ws_bytes(n, rows)) {
    set_error("exclusive_scan: workspace too small (%zu < %zu)", ws_bytes, scan_ws_bytes(n, rows));
    return ISOB200_ERR_WORKSPACE;
  }
  int* tile_sums = (int*)ws;
  scan_reduce_kernel<<<dim3(tiles, rows), SCAN_THREADS, 0, stream>>>(in, n, in_stride, tiles,
                                                                     tile_sums);
  ISO_CHECK_LAUNCH("scan_reduce");
  scan_spine_kernel<<<rows, SCAN_THREADS, 0, stream>>>(tile_sums, tiles);
  ISO_CHECK_LAUNCH("scan_spine");
  scan_downsweep_kernel<<<dim3(tiles, rows), SCAN_THREADS, 0, stream>>>(
      in, out, n, in_stride, out_stride, tiles, tile_sums);
  ISO_CHECK_LAUNCH("scan_downsweep");
  return ISOB200_OK;
}

}  // namespace isob200

extern "C" {

size_t isob200_exclusive_scan_ws_bytes(int n, int rows) { return isob200::scan_ws_bytes(n, rows); }

int isob200_exclusive_scan_i32(const int* in, int* out, int n, int rows, long long in_stride,
                               long long out_stride, void* ws, size_t ws_bytes, void* stream) {
  ISO_CHECK_ARG(n >= 0 && rows >= 0, "exclusive_scan: negative size");
  ISO_CHECK_ARG(n == 0 || rows == 0 || (in && out), "exclusive_scan: null pointer");
  return isob200::exclusive_scan_i32(in, out, n, rows, in_stride, out_stride, ws, ws_bytes,
                                     (cudaStream_t)stream);
}

}  // extern "C"
