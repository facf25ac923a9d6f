#pragma once
#include "common.cuh"
namespace isob200 {
size_t scan_ws_bytes(int n, int rows);
// out[row][i] = sum_{j<i} in[row][j]; rows are `*_stride` elements apart. `out` may alias `in`.
int exclusive_scan_i32(const int* in, int* out, int n, int rows, long long in_stride,
                       long long out_stride, void* ws, size_t ws_bytes, cudaStream_t stream);

// ---- scan in the tail of the kernel that produced the counts ---------------------------------------------------
// For count tables of up to LAST_CTA_SCAN_MAX entries the three scan launches cost more than the scan (one CTA walks
// the table with 16-byte loads: ~4 us at 16 k entries; at 60 k entries -- the digit tables of the FRNN radix build --
// it was measured slower than the three launches, 0.10 -> 0.19 ms per build, and is not used there): every CTA
// (256 threads) of the producing kernel calls this once its own counts are in global memory; the CTA that draws the
// last ticket turns cnt[0, n) into exclusive offsets off[0, n) and writes the grand total to total_a / total_b
// (either may be null).  `ticket` is a zeroed int; cnt / off are 16-byte aligned; `off` may be `cnt` (in place:
// a thread re-reads only its own segment before overwriting it).
constexpr int LAST_CTA_SCAN_MAX = 1 << 14;

#ifdef __CUDACC__
__device__ __forceinline__ void last_cta_exclusive_scan(int* ticket, int n_ctas, const int* cnt, int* off, int n,
                                                        int* total_a, int* total_b) {
  __shared__ int s_part[8];
  __shared__ bool s_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(ticket, 1) == n_ctas - 1;
  __syncthreads();
  if (!s_last) return;
  const int per = (((n + 255) / 256) + 3) & ~3;            // a thread's segment: whole int4s
  const int lo = min((int)threadIdx.x * per, n), hi = min(lo + per, n);
  int sum = 0;
  {
    int i = lo;
#pragma unroll 4
    for (; i + 4 <= hi; i += 4) {                          // the other CTAs' writes / atomics live in L2: ld.cg
      const int4 v = __ldcg(reinterpret_cast<const int4*>(cnt + i));
      sum += v.x + v.y + v.z + v.w;
    }
    for (; i < hi; ++i) sum += __ldcg(cnt + i);
  }
  // block-wide exclusive scan of the 256 segment sums: shuffle scan per warp, then over the 8 warp totals
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int inc = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) s_part[warp] = inc;
  __syncthreads();
  int base = 0, grand = 0;
#pragma unroll
  for (int w = 0; w < 8; ++w) {
    const int t = s_part[w];
    if (w < warp) base += t;
    grand += t;
  }
  if (threadIdx.x == 0) {
    if (total_a) total_a[0] = grand;
    if (total_b) total_b[0] = grand;
  }
  int acc = base + inc - sum;
  {
    int i = lo;
#pragma unroll 4
    for (; i + 4 <= hi; i += 4) {
      const int4 v = __ldcg(reinterpret_cast<const int4*>(cnt + i));
      int4 o;
      o.x = acc; o.y = acc + v.x; o.z = o.y + v.y; o.w = o.z + v.z;
      acc = o.w + v.w;
      *reinterpret_cast<int4*>(off + i) = o;
    }
    for (; i < hi; ++i) { off[i] = acc; acc += __ldcg(cnt + i); }
  }
}
#endif
}  // namespace isob200
