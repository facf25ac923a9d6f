// DSS elliptical-splat rasteriser, forward, for sm_100a.
//
// Replaces DSS/csrc/rasterize_points.cu: RasterizePoints{Naive,Coarse,Fine}CudaKernel
// (:131-212, :293-432, :506-597) behind DSS._C.splat_points (rasterize_points.h:461-525).
//
// Reference design: a (N,B,B,M) bin_points matrix with M = max(10000, Pmax) int32 slots per
// 32x32-pixel bin (2.46 GB for 8 views x 300 k points), filled by a 64-block bitmask kernel, then
// ONE THREAD PER PIXEL walking all M slots with a 150-entry Pix list in local memory.
//
// B200 design (HBM / L2-bound integer + fp32 work, no dense contraction):
//   1. splat_tile_count : one thread per point -> exact range of 16x16-pixel tiles its radii box
//      can touch -> per-tile counters (integer atomics: order independent).
//   2. exclusive scan of the counters (scan.cu) -> tile offsets (+ total, read back once to size
//      the record buffer).
//   3. splat_tile_fill  : each (tile, point) pair becomes a packed 48-byte record
//      {px,py,pz,a | b,c,cutoff,rx | ry,id,-,-} in the tile's contiguous slice (cursor atomics;
//      slot order inside a tile is irrelevant: the per-pixel selection below is a total order).
//   4. splat_raster<K>  : one CTA per tile, 8 warps, each warp owns an 8x4 pixel block.  The
//      tile's record slice is streamed through shared memory in 128-record chunks by the TMA
//      engine (cp.async.bulk + mbarrier, double buffered).  Per 32 records a warp does a
//      one-lane-per-record box test against its pixel block, ballots the survivors, and only
//      those are evaluated per pixel (exact reference predicate) and inserted into a K-entry
//      (z, id, Q) list held in REGISTERS, sorted by (z, id).  Outputs idx/zbuf/qvalue/occ are
//      written once, vectorised, -1 padding included (no fill pass).  (The default raster kernel is v2 below:
//      record-centric hit generation into per-pixel shared-memory lists, then one thread per pixel; its
//      epilogue optionally computes the renderer's RGBA blend and the per-point visibility of the backward
//      while the pixel's K entries are still in registers -- isob200_splat_forward_fused.)
// Result = reference on tie-free depth: the K covering points with the smallest z, ascending,
// cut at z - z0 > depth_merging_thres (:203-206); equal z is ordered by smaller point id (the
// reference keeps whichever its atomics inserted first).
// fp32 expressions that decide coverage are spelled with intrinsics in the exact order nvcc
// contracts the reference's (SASS of RasterizePointsFineCudaKernel):
//   xf = (float(2i)+1)/S - 1 ; dx = xf - px ; Q = fma(c*dy, dy, fma(a*dx, dx, (b*dx)*dy)).
#include "common.cuh"
#include "scan.cuh"
#include <float.h>
#include <limits.h>

namespace isob200 {

constexpr int TILE = 16;             // pixels per tile side
constexpr int REC_F4 = 3;            // float4s per record (48 B)
constexpr int CHUNK = 128;           // records per TMA chunk (6 KB)
constexpr int STAGES = 2;

__device__ __forceinline__ float pix_to_ndc(int i, float fS) {
  // rasterization_utils.cuh:8-11: -1 + (2*i + 1.0f) / S
  return __fadd_rn(__fdiv_rn(__fadd_rn((float)(2 * i), 1.0f), fS), -1.0f);
}

// Conservative-then-trimmed range of pixel indices i in [0,S) with |ndc(i) - p| <= r.
// Returns lo > hi when empty.  Never a subset of the exact set (estimate error << 1 pixel and
// the bounds start one pixel outside).
__device__ __forceinline__ void pixel_range(float p, float r, int S, float fS, int& lo, int& hi) {
  const float e_lo = (p - r + 1.0f) * fS * 0.5f - 0.5f;
  const float e_hi = (p + r + 1.0f) * fS * 0.5f - 0.5f;
  // clamp in float first: huge radii / far-away points must not overflow the int conversion
  lo = (int)fmaxf(ceilf(e_lo) - 1.0f, 0.0f);
  hi = (int)fminf(floorf(e_hi) + 1.0f, (float)(S - 1));
  if (!(e_lo <= (float)S) || !(e_hi >= -1.0f)) { lo = 1; hi = 0; return; }  // also rejects NaN
  if (lo <= hi && fabsf(__fsub_rn(pix_to_ndc(lo, fS), p)) > r) ++lo;
  if (lo <= hi && fabsf(__fsub_rn(pix_to_ndc(hi, fS), p)) > r) --hi;
}

struct TileRect { int tx0, tx1, ty0, ty1; };

// Tiles are indexed in OUTPUT image coordinates (col = S-1-xi, row = S-1-yi, the flip of
// rasterize_points.cu:160-161 / :577-578) so that the raster kernel's stores are coalesced.
__device__ __forceinline__ bool tile_rect_of(float px, float py, float pz, float rx, float ry, int S, float fS,
                                             TileRect& t) {
  if (!(pz >= 0.0f)) return false;                       // behind the camera (:87-88)
  int xl, xh, yl, yh;
  pixel_range(px, rx, S, fS, xl, xh);
  pixel_range(py, ry, S, fS, yl, yh);
  if (xl > xh || yl > yh) return false;
  t.tx0 = (S - 1 - xh) / TILE; t.tx1 = (S - 1 - xl) / TILE;
  t.ty0 = (S - 1 - yh) / TILE; t.ty1 = (S - 1 - yl) / TILE;
  return true;
}

__device__ __forceinline__ bool point_tile_rect(const float* __restrict__ points,
                                                const float* __restrict__ radii, long long p, int S,
                                                float fS, TileRect& t) {
  const float px = points[3 * p], py = points[3 * p + 1], pz = points[3 * p + 2];
  if (!(pz >= 0.0f)) return false;                       // behind the camera (:87-88)
  const float rx = radii[2 * p], ry = radii[2 * p + 1];
  int xl, xh, yl, yh;
  pixel_range(px, rx, S, fS, xl, xh);
  pixel_range(py, ry, S, fS, yl, yh);
  if (xl > xh || yl > yh) return false;
  t.tx0 = (S - 1 - xh) / TILE; t.tx1 = (S - 1 - xl) / TILE;
  t.ty0 = (S - 1 - yh) / TILE; t.ty1 = (S - 1 - yl) / TILE;
  return true;
}

__global__ void __launch_bounds__(256)
splat_tile_count_kernel(const float* __restrict__ points, const float* __restrict__ radii,
                        const int64_t* __restrict__ first_idx, const int64_t* __restrict__ num_points,
                        int S, int T, int* __restrict__ tile_cnt) {
  const int n = blockIdx.y;
  const long long first = first_idx[n];
  const long long num = num_points[n];
  const float fS = (float)S;
  int* cnt = tile_cnt + (size_t)n * T * T;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < num;
       i += (long long)gridDim.x * blockDim.x) {
    TileRect t;
    if (!point_tile_rect(points, radii, first + i, S, fS, t)) continue;
    for (int ty = t.ty0; ty <= t.ty1; ++ty)
      for (int tx = t.tx0; tx <= t.tx1; ++tx) atomicAdd(&cnt[ty * T + tx], 1);
  }
}

// total[0] = number of (tile, point) records = off[last] + cnt[last]
__global__ void splat_total_kernel(const int* __restrict__ cnt, const int* __restrict__ off, int ntiles,
                                   int* __restrict__ total) {
  if (threadIdx.x == 0 && blockIdx.x == 0) total[0] = ntiles > 0 ? off[ntiles - 1] + cnt[ntiles - 1] : 0;
}

__global__ void __launch_bounds__(256)
splat_tile_fill_kernel(const float* __restrict__ points, const float* __restrict__ ellipse,
                       const float* __restrict__ cutoff, const float* __restrict__ radii,
                       const int64_t* __restrict__ first_idx, const int64_t* __restrict__ num_points,
                       int S, int T, const int* __restrict__ tile_off, int* __restrict__ tile_cur,
                       long long capacity, float4* __restrict__ recs) {
  const int n = blockIdx.y;
  const long long first = first_idx[n];
  const long long num = num_points[n];
  const float fS = (float)S;
  const int* off = tile_off + (size_t)n * T * T;
  int* cur = tile_cur + (size_t)n * T * T;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < num;
       i += (long long)gridDim.x * blockDim.x) {
    const long long p = first + i;
    TileRect t;
    if (!point_tile_rect(points, radii, p, S, fS, t)) continue;
    const float4 r0 = make_float4(points[3 * p], points[3 * p + 1], points[3 * p + 2], ellipse[3 * p]);
    const float4 r1 = make_float4(ellipse[3 * p + 1], ellipse[3 * p + 2], cutoff[p], radii[2 * p]);
    const float4 r2 = make_float4(radii[2 * p + 1], __int_as_float((int)p), 0.f, 0.f);
    for (int ty = t.ty0; ty <= t.ty1; ++ty)
      for (int tx = t.tx0; tx <= t.tx1; ++tx) {
        const int tile = ty * T + tx;
        const long long slot = (long long)off[tile] + atomicAdd(&cur[tile], 1);
        if (slot < capacity) {
          recs[slot * REC_F4 + 0] = r0;
          recs[slot * REC_F4 + 1] = r1;
          recs[slot * REC_F4 + 2] = r2;
        }
      }
  }
}

// ---- privatised form of the count kernel ------------------------------------------------------------------------
// The plain count kernel issues one global atomic per (tile, point) pair: 3.4 M atomics on 8 192 counters at BASELINE
// config 4, ~415 per address, and the L2 atomic unit serialises per address.  Here a CTA takes a contiguous chunk of
// >= 4 096 points of ONE view, bins it into a shared-memory histogram of that view's tiles and adds the non-zero
// bins to the global counters: 0.075 -> 0.05 ms (chunks of 8 192 left half the SMs without a CTA: 0.059 ms for
// the whole binning entry against 0.048; 2 048 is the same as 4 096).  (The same idea for the fill kernel -- reserve a CTA's range of
// every tile's slice with one atomic, hand the slots out from shared memory in a second pass -- was measured and
// dropped: two passes on 300 CTAs write the 163 MB of records slower than one pass on 1 184, 0.12 -> 0.20 ms.)
constexpr int PRIV_MAX_TILES = 4096;      // tiles per view the shared-memory histogram holds (16 KB)
constexpr int PRIV_CHUNK = 4096;          // points per CTA

__global__ void __launch_bounds__(256)
splat_tile_count_priv_kernel(const float* __restrict__ points, const float* __restrict__ radii,
                             const int64_t* __restrict__ first_idx, const int64_t* __restrict__ num_points,
                             int S, int T, int* __restrict__ tile_cnt, int* __restrict__ ticket,
                             int* __restrict__ tile_off, int* __restrict__ total, int* __restrict__ total_out) {
  extern __shared__ int s_hist[];          // [T*T]
  const int n = blockIdx.y, nt = T * T;
  const long long first = first_idx[n], num = num_points[n];
  const long long chunk = (num + gridDim.x - 1) / gridDim.x;
  const long long b = (long long)blockIdx.x * chunk, e = min(num, b + chunk);
  if (b < e) {
    for (int i = threadIdx.x; i < nt; i += blockDim.x) s_hist[i] = 0;
    __syncthreads();
    const float fS = (float)S;
    // four points' coordinates and radii are loaded before any is used (few CTAs, one dependent load chain per
    // thread otherwise: the kernel ran at the memory latency)
    for (long long i0 = b + threadIdx.x; i0 < e; i0 += 4 * blockDim.x) {
      float px[4], py[4], pz[4], rx[4], ry[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const long long i = i0 + (long long)u * blockDim.x;
        const long long p = first + (i < e ? i : i0);
        px[u] = points[3 * p]; py[u] = points[3 * p + 1]; pz[u] = i < e ? points[3 * p + 2] : -1.0f;
        rx[u] = radii[2 * p]; ry[u] = radii[2 * p + 1];
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        TileRect t;
        if (!tile_rect_of(px[u], py[u], pz[u], rx[u], ry[u], S, fS, t)) continue;
        for (int ty = t.ty0; ty <= t.ty1; ++ty)
          for (int tx = t.tx0; tx <= t.tx1; ++tx) atomicAdd(&s_hist[ty * T + tx], 1);
      }
    }
    __syncthreads();
    int* cnt = tile_cnt + (size_t)n * nt;
    for (int i = threadIdx.x; i < nt; i += blockDim.x)
      if (s_hist[i]) atomicAdd(&cnt[i], s_hist[i]);
  }
  // the last CTA to finish turns the counts of all views into offsets and the record total: no scan launches, no
  // separate total kernel, no device-to-device copy behind this kernel
  if (ticket)
    last_cta_exclusive_scan(ticket, (int)(gridDim.x * gridDim.y), tile_cnt, tile_off, nt * (int)gridDim.y, total,
                            total_out);
}

// ---- TMA / mbarrier helpers (cp.async.bulk: SASS UBLKCP) ----
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned phase) {
  unsigned done;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(phase)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ void tma_bulk_g2s(void* smem_dst, const void* gmem_src, unsigned bytes,
                                             unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// (z, id) total order of the per-pixel selection
__device__ __forceinline__ bool zid_less(float z, int id, float z2, int id2) {
  return (z < z2) || (z == z2 && id < id2);
}

// Per-pixel candidate list in registers: the reference's algorithm (unsorted K slots + running
// maximum, rasterize_points.cu:99-123) with the order made total by (z, id); ONE sorting network at
// the end instead of a sorted insertion per hit -- the warp executes the list update whenever any of
// its 32 pixels is hit, so its instruction count is what bounds the kernel.
template <int K>
struct PixList {
  float z[K];
  int id[K];
  float q[K];
  int cnt;
  float maxz;
  int maxid, maxpos;
  __device__ __forceinline__ void clear() {
#pragma unroll
    for (int k = 0; k < K; ++k) { z[k] = FLT_MAX; id[k] = INT_MAX; q[k] = 0.f; }
    cnt = 0; maxz = -FLT_MAX; maxid = -1; maxpos = 0;
  }
  __device__ __forceinline__ void insert(float cz, int cid, float cq) {
    if (cnt < K) {
#pragma unroll
      for (int k = 0; k < K; ++k)
        if (k == cnt) { z[k] = cz; id[k] = cid; q[k] = cq; }
      if (zid_less(maxz, maxid, cz, cid)) { maxz = cz; maxid = cid; maxpos = cnt; }
      ++cnt;
    } else if (zid_less(cz, cid, maxz, maxid)) {
#pragma unroll
      for (int k = 0; k < K; ++k)
        if (k == maxpos) { z[k] = cz; id[k] = cid; q[k] = cq; }
      maxz = z[0]; maxid = id[0]; maxpos = 0;
#pragma unroll
      for (int k = 1; k < K; ++k)
        if (zid_less(maxz, maxid, z[k], id[k])) { maxz = z[k]; maxid = id[k]; maxpos = k; }
    }
  }
  __device__ __forceinline__ void cswap(int a, int b) {
    const bool sw = zid_less(z[b], id[b], z[a], id[a]);
    const float tz = sw ? z[b] : z[a], uz = sw ? z[a] : z[b];
    const int ti = sw ? id[b] : id[a], ui = sw ? id[a] : id[b];
    const float tq = sw ? q[b] : q[a], uq = sw ? q[a] : q[b];
    z[a] = tz; z[b] = uz; id[a] = ti; id[b] = ui; q[a] = tq; q[b] = uq;
  }
  // ascending (z, id); empty slots (FLT_MAX, INT_MAX) sink to the end.  Odd-even transposition network.
  __device__ __forceinline__ void sort() {
#pragma unroll
    for (int r = 0; r < K; ++r) {
#pragma unroll
      for (int k = (r & 1); k + 1 < K; k += 2) cswap(k, k + 1);
    }
  }
};

// The unsorted phase of the list lives in SHARED memory, one column per thread ([k][tid]: a thread's
// slot k sits in bank tid % 32 whatever k is, so dynamically indexed appends are conflict-free and cost
// three STS instead of a K-way chain of predicated register moves).  Registers keep only the count and
// the running maximum; the K slots are pulled into registers once, for the final sorting network.
template <int K>
struct SmemList {
  float* z;      // [K][256]
  int* id;
  float* q;
  int cnt, maxid, maxpos;
  float maxz;
  __device__ __forceinline__ void init(float* base) {
    z = base + threadIdx.x;
    id = reinterpret_cast<int*>(base + K * 256) + threadIdx.x;
    q = base + 2 * K * 256 + threadIdx.x;
    cnt = 0; maxz = -FLT_MAX; maxid = -1; maxpos = 0;
  }
  __device__ __forceinline__ void insert(float cz, int cid, float cq) {
    if (cnt < K) {
      z[cnt * 256] = cz; id[cnt * 256] = cid; q[cnt * 256] = cq;
      if (zid_less(maxz, maxid, cz, cid)) { maxz = cz; maxid = cid; maxpos = cnt; }
      ++cnt;
    } else if (zid_less(cz, cid, maxz, maxid)) {
      z[maxpos * 256] = cz; id[maxpos * 256] = cid; q[maxpos * 256] = cq;
      maxz = z[0]; maxid = id[0]; maxpos = 0;
#pragma unroll
      for (int k = 1; k < K; ++k) {
        const float kz = z[k * 256];
        const int ki = id[k * 256];
        if (zid_less(maxz, maxid, kz, ki)) { maxz = kz; maxid = ki; maxpos = k; }
      }
    }
  }
  __device__ __forceinline__ void drain(PixList<K>& L) const {
    L.clear();
#pragma unroll
    for (int k = 0; k < K; ++k)
      if (k < cnt) { L.z[k] = z[k * 256]; L.id[k] = id[k * 256]; L.q[k] = q[k * 256]; }
    L.cnt = cnt; L.maxz = maxz; L.maxid = maxid; L.maxpos = maxpos;
  }
};

template <int K>
__global__ void __launch_bounds__(256)
splat_raster_kernel(const float4* __restrict__ recs, const int* __restrict__ tile_off,
                    const int* __restrict__ tile_cnt, int S, int T, float depth_merging_thres,
                    int occ_inclusive, int* __restrict__ out_idx, float* __restrict__ out_z,
                    float* __restrict__ out_q, float* __restrict__ out_occ) {
  __shared__ __align__(128) float4 buf[STAGES][CHUNK * REC_F4];
  __shared__ __align__(8) unsigned long long bar[STAGES];

  const int tile = blockIdx.x;                 // n*T*T + ty*T + tx
  const int n = tile / (T * T);
  const int tr = tile - n * T * T;
  const int ty = tr / T, tx = tr - ty * T;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c0 = tx * TILE + (warp & 1) * 8;   // first output column of this warp's 8x4 block
  const int r0 = ty * TILE + (warp >> 1) * 4;
  const int col = c0 + (lane & 7), row = r0 + (lane >> 3);
  const bool in_img = col < S && row < S;
  const float fS = (float)S;
  const int xi = S - 1 - col, yi = S - 1 - row;   // NDC pixel indices (may be < 0 outside the image)
  const float xf = pix_to_ndc(xi, fS), yf = pix_to_ndc(yi, fS);
  // NDC box of the warp's block, grown by half a pixel: a record whose radii box misses it
  // cannot pass the exact per-pixel test for any of the 32 pixels
  const float half = 1.0f / fS;
  const float bx_lo = pix_to_ndc(S - 1 - (c0 + 7), fS) - half, bx_hi = pix_to_ndc(S - 1 - c0, fS) + half;
  const float by_lo = pix_to_ndc(S - 1 - (r0 + 3), fS) - half, by_hi = pix_to_ndc(S - 1 - r0, fS) + half;

  const int nrec = tile_cnt[tile];
  const long long base = tile_off[tile];
  const int nchunks = (nrec + CHUNK - 1) / CHUNK;

  if (threadIdx.x == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int c = 0; c < STAGES && c < nchunks; ++c) {
      const int cnt = min(CHUNK, nrec - c * CHUNK);
      const unsigned bytes = (unsigned)cnt * REC_F4 * 16u;
      mbar_expect_tx(&bar[c], bytes);
      tma_bulk_g2s(&buf[c][0], recs + (base + (long long)c * CHUNK) * REC_F4, bytes, &bar[c]);
    }
  }

  extern __shared__ __align__(16) float list_smem[];   // 3 * K * 256 words
  SmemList<K> SL;
  SL.init(list_smem);

  for (int c = 0; c < nchunks; ++c) {
    const int st = c & 1;
    mbar_wait(&bar[st], (unsigned)((c >> 1) & 1));
    const int cnt = min(CHUNK, nrec - c * CHUNK);
    const float4* rb = buf[st];
    for (int j0 = 0; j0 < cnt; j0 += 32) {
      const int j = j0 + lane;
      bool hit = false;
      if (j < cnt) {
        const float4 a0 = rb[j * REC_F4 + 0];
        const float rx = rb[j * REC_F4 + 1].w;
        const float ry = rb[j * REC_F4 + 2].x;
        hit = !(a0.x - rx > bx_hi) && !(a0.x + rx < bx_lo) && !(a0.y - ry > by_hi) && !(a0.y + ry < by_lo);
      }
      unsigned m = __ballot_sync(0xffffffffu, hit);
      while (m) {
        const int s = j0 + __ffs(m) - 1;
        m &= m - 1;
        const float4 a0 = rb[s * REC_F4 + 0];   // px py pz a   (smem broadcast)
        const float4 a1 = rb[s * REC_F4 + 1];   // b  c  cutoff rx
        const float4 a2 = rb[s * REC_F4 + 2];   // ry id
        // CheckPixelInsidePoint, rasterize_points.cu:64-98
        const float dx = __fsub_rn(xf, a0.x);
        const float dy = __fsub_rn(yf, a0.y);
        if (fabsf(dx) > a1.w || fabsf(dy) > a2.x) continue;
        const float q = __fmaf_rn(__fmul_rn(a1.y, dy), dy,
                                  __fmaf_rn(__fmul_rn(a0.w, dx), dx, __fmul_rn(__fmul_rn(a1.x, dx), dy)));
        if (q > a1.z) continue;
        SL.insert(a0.z, __float_as_int(a2.y), q);
      }
    }
    __syncthreads();   // every warp is done with this stage
    if (threadIdx.x == 0 && c + STAGES < nchunks) {
      const int c2 = c + STAGES;
      const int cnt2 = min(CHUNK, nrec - c2 * CHUNK);
      const unsigned bytes = (unsigned)cnt2 * REC_F4 * 16u;
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_expect_tx(&bar[st], bytes);
      tma_bulk_g2s(&buf[st][0], recs + (base + (long long)c2 * CHUNK) * REC_F4, bytes, &bar[st]);
    }
  }

  if (!in_img) return;
  // epilogue: sort, depth-merge cut (:203-206), occupancy (:196 naive >=, :581 fine >), -1 padding
  PixList<K> L;
  SL.drain(L);
  L.sort();
  const int size = L.cnt;
  const float zmax = size > 0 ? L.maxz : -1000.0f;
  const float z0 = L.z[0];
  bool alive = true;
  int oi[K];
  float oz[K], oq[K];
#pragma unroll
  for (int k = 0; k < K; ++k) {
    alive = alive && (k < size) && !(__fsub_rn(L.z[k], z0) > depth_merging_thres);
    oi[k] = alive ? L.id[k] : -1;
    oz[k] = alive ? L.z[k] : -1.0f;
    oq[k] = alive ? L.q[k] : -1.0f;
  }
  const size_t pix = ((size_t)n * S + row) * S + col;
  out_occ[pix] = (size > 0 && (occ_inclusive ? (zmax >= 0.0f) : (zmax > 0.0f))) ? 1.0f : 0.0f;
  int* pi = out_idx + pix * K;
  float* pz = out_z + pix * K;
  float* pq = out_q + pix * K;
  if constexpr (K % 4 == 0) {
#pragma unroll
    for (int k = 0; k < K; k += 4) {
      *reinterpret_cast<int4*>(pi + k) = make_int4(oi[k], oi[k + 1], oi[k + 2], oi[k + 3]);
      *reinterpret_cast<float4*>(pz + k) = make_float4(oz[k], oz[k + 1], oz[k + 2], oz[k + 3]);
      *reinterpret_cast<float4*>(pq + k) = make_float4(oq[k], oq[k + 1], oq[k + 2], oq[k + 3]);
    }
  } else if constexpr (K % 2 == 0) {
#pragma unroll
    for (int k = 0; k < K; k += 2) {
      *reinterpret_cast<int2*>(pi + k) = make_int2(oi[k], oi[k + 1]);
      *reinterpret_cast<float2*>(pz + k) = make_float2(oz[k], oz[k + 1]);
      *reinterpret_cast<float2*>(pq + k) = make_float2(oq[k], oq[k + 1]);
    }
  } else {
#pragma unroll
    for (int k = 0; k < K; ++k) { pi[k] = oi[k]; pz[k] = oz[k]; pq[k] = oq[k]; }
  }
}

// ---------------------------------------------------------------------------------------------
// Raster v2: record-centric hit generation.
// v1 above evaluates every surviving record on all 32 pixels of a warp's 8x4 block although a
// sigma = 1.5 px splat covers ~2 of them: ~7 % of the issued lanes do useful work, and the kernel is
// issue-bound.  Here the roles are swapped for the collection phase:
//   phase 1  one THREAD per record sweeps the record's own (exactly trimmed) pixel box inside the
//            tile and appends each hit (z, id, Q) to that pixel's column in shared memory
//            (slot = atomicAdd on a per-pixel shared counter; C slots per pixel);
//   phase 2  one thread per pixel keeps the K smallest (z, id) of its <= C entries in place
//            (replace-max), pulls them into registers and sorts them with the network of v1.
// The arrival order inside a column depends on scheduling, the selected set and its order do not:
// (z, id) is a total order.  A pixel that receives more than C hits (dense overdraw) is redone by
// the v1 per-pixel algorithm in a second pass over the tile's records, so any input is handled.
// ---------------------------------------------------------------------------------------------
constexpr int CHUNK2 = 256;          // records per TMA chunk in v2: one per thread (12 KB)

// Optional work fused into the raster epilogue, where a pixel's K (id, Q) entries are still in registers:
//  * RGBA blend (DSS/core/renderer.py:53-78 + pytorch3d's NormWeightedCompositor): img = [sum_k w_k f[id_k] /
//    max(sum_k w_k, eps), occ], w_k = exp(-Q_k / 2) * scaler[id_k] -- the arithmetic of splat_blend_kernel
//    (splat_bwd.cu) in the same order, without its second read of idx / qvalue (64 B per pixel);
//  * per-point visibility for the backward (rasterizer.py:851-857): every id of a pixel whose first slot is taken.
struct RasterEpi {
  const float* scaler;   // (P) or null
  const float* feat;     // (P, feat_stride), C <= 4 channels used; null = no blend
  int feat_stride, C;
  float eps;
  float* img;            // (N,S,S,C+1)
  float* weights;        // (N,S,S,K) or null
  unsigned char* visible;   // (P) or null; zeroed by the caller
};

template <int K, int C, int STG>
__global__ void __launch_bounds__(256)
splat_raster_v2_kernel(const float4* __restrict__ recs, const int* __restrict__ tile_off,
                       const int* __restrict__ tile_cnt, int S, int T, float depth_merging_thres,
                       int occ_inclusive, int* __restrict__ out_idx, float* __restrict__ out_z,
                       float* __restrict__ out_q, float* __restrict__ out_occ, const RasterEpi epi) {
  __shared__ __align__(128) float4 buf[STG][CHUNK2 * REC_F4];
  __shared__ __align__(8) unsigned long long bar[STG];
  __shared__ int pcnt[256];
  __shared__ float ndcx[TILE], ndcy[TILE];
  extern __shared__ __align__(16) float list_smem[];   // [3][C][256]: z | id | q columns per pixel
  float* lz = list_smem;
  int* lid = reinterpret_cast<int*>(list_smem + C * 256);
  float* lq = list_smem + 2 * C * 256;

  const int tile = blockIdx.x;
  const int n = tile / (T * T);
  const int tr = tile - n * T * T;
  const int ty = tr / T, tx = tr - ty * T;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float fS = (float)S;
  const int nrec = tile_cnt[tile];
  const long long base = tile_off[tile];
  const int nchunks = (nrec + CHUNK2 - 1) / CHUNK2;
  // number of tile columns / rows that lie inside the image
  const int ncol_img = min(TILE, S - tx * TILE), nrow_img = min(TILE, S - ty * TILE);

  pcnt[tid] = 0;
  if (tid < TILE) ndcx[tid] = pix_to_ndc(S - 1 - (tx * TILE + tid), fS);
  else if (tid < 2 * TILE) ndcy[tid - TILE] = pix_to_ndc(S - 1 - (ty * TILE + tid - TILE), fS);
  if (tid == 0) {
#pragma unroll
    for (int g = 0; g < STG; ++g) mbar_init(&bar[g], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  int gc = 0;   // chunks streamed so far over both passes: stage = gc % STG, parity = (gc / STG) & 1
  auto issue = [&](int c, int g) {
    const int cnt = min(CHUNK2, nrec - c * CHUNK2);
    const unsigned bytes = (unsigned)cnt * REC_F4 * 16u;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    mbar_expect_tx(&bar[g % STG], bytes);
    tma_bulk_g2s(&buf[g % STG][0], recs + (base + (long long)c * CHUNK2) * REC_F4, bytes, &bar[g % STG]);
  };

  // ---- phase 1: hit generation, one thread per record ----
  if (tid == 0)
    for (int c = 0; c < STG && c < nchunks; ++c) issue(c, gc + c);
  for (int c = 0; c < nchunks; ++c, ++gc) {
    mbar_wait(&bar[gc % STG], (unsigned)((gc / STG) & 1));
    const int cnt = min(CHUNK2, nrec - c * CHUNK2);
    if (tid < cnt) {
      const float4* rb = buf[gc % STG] + tid * REC_F4;
      const float4 a0 = rb[0], a1 = rb[1], a2 = rb[2];
      int xl, xh, yl, yh;
      pixel_range(a0.x, a1.w, S, fS, xl, xh);
      pixel_range(a0.y, a2.x, S, fS, yl, yh);
      // to tile-local output coordinates (col = S-1-xi)
      const int cx0 = max(S - 1 - xh - tx * TILE, 0), cx1 = min(S - 1 - xl - tx * TILE, ncol_img - 1);
      const int cy0 = max(S - 1 - yh - ty * TILE, 0), cy1 = min(S - 1 - yl - ty * TILE, nrow_img - 1);
      const int id = __float_as_int(a2.y);
      for (int yy = cy0; yy <= cy1; ++yy) {
        const float dy = __fsub_rn(ndcy[yy], a0.y);
        if (fabsf(dy) > a2.x) continue;
        const float cdy = __fmul_rn(a1.y, dy);
        for (int xx = cx0; xx <= cx1; ++xx) {
          const float dx = __fsub_rn(ndcx[xx], a0.x);
          if (fabsf(dx) > a1.w) continue;
          const float q = __fmaf_rn(cdy, dy, __fmaf_rn(__fmul_rn(a0.w, dx), dx, __fmul_rn(__fmul_rn(a1.x, dx), dy)));
          if (q > a1.z) continue;
          const int pix = yy * TILE + xx;
          const int slot = atomicAdd(&pcnt[pix], 1);
          if (slot < C) { lz[slot * 256 + pix] = a0.z; lid[slot * 256 + pix] = id; lq[slot * 256 + pix] = q; }
        }
      }
    }
    __syncthreads();
    if (tid == 0 && c + STG < nchunks) issue(c + STG, gc + STG);
  }

  // ---- phase 2: per-pixel selection ----
  const int col = tx * TILE + (tid & 15), row = ty * TILE + (tid >> 4);
  const bool in_img = col < S && row < S;
  const int nhit = pcnt[tid];
  const bool overflow = nhit > C;
  PixList<K> L;
  L.clear();
  if (__syncthreads_or(overflow)) {
    // dense overdraw somewhere in this tile: the overflowing pixels are redone exactly by the v1
    // per-pixel algorithm (warp-level box cull + replace-max list in the first K slots of the column)
    SmemList<K> SL;
    SL.z = lz + tid; SL.id = lid + tid; SL.q = lq + tid;
    SL.cnt = 0; SL.maxz = -FLT_MAX; SL.maxid = -1; SL.maxpos = 0;
    const float xf = ndcx[tid & 15], yf = ndcy[tid >> 4];
    const unsigned need = __ballot_sync(0xffffffffu, overflow);
    if (tid == 0)
      for (int c = 0; c < STG && c < nchunks; ++c) issue(c, gc + c);
    for (int c = 0; c < nchunks; ++c, ++gc) {
      mbar_wait(&bar[gc % STG], (unsigned)((gc / STG) & 1));
      const int cnt = min(CHUNK2, nrec - c * CHUNK2);
      if (need) {
        const float4* rb = buf[gc % STG];
        for (int s = 0; s < cnt; ++s) {
          const float4 a0 = rb[s * REC_F4 + 0], a1 = rb[s * REC_F4 + 1], a2 = rb[s * REC_F4 + 2];
          const float dx = __fsub_rn(xf, a0.x), dy = __fsub_rn(yf, a0.y);
          if (!overflow || fabsf(dx) > a1.w || fabsf(dy) > a2.x) continue;
          const float q = __fmaf_rn(__fmul_rn(a1.y, dy), dy,
                                    __fmaf_rn(__fmul_rn(a0.w, dx), dx, __fmul_rn(__fmul_rn(a1.x, dx), dy)));
          if (q > a1.z) continue;
          SL.insert(a0.z, __float_as_int(a2.y), q);
        }
      }
      __syncthreads();
      if (tid == 0 && c + STG < nchunks) issue(c + STG, gc + STG);
    }
    if (overflow) SL.drain(L);
  }
  if (!in_img) return;
  if (!overflow) {
    float* cz = lz + tid; int* ci = lid + tid; float* cq = lq + tid;
    if (nhit > K) {
      // keep the K smallest (z, id) in slots [0, K): replace-max over the remaining entries
      float maxz = cz[0]; int maxid = ci[0], maxpos = 0;
#pragma unroll
      for (int k = 1; k < K; ++k)
        if (zid_less(maxz, maxid, cz[k * 256], ci[k * 256])) { maxz = cz[k * 256]; maxid = ci[k * 256]; maxpos = k; }
      for (int e = K; e < nhit; ++e) {
        const float ez = cz[e * 256];
        const int ei = ci[e * 256];
        if (!zid_less(ez, ei, maxz, maxid)) continue;
        cz[maxpos * 256] = ez; ci[maxpos * 256] = ei; cq[maxpos * 256] = cq[e * 256];
        maxz = cz[0]; maxid = ci[0]; maxpos = 0;
#pragma unroll
        for (int k = 1; k < K; ++k)
          if (zid_less(maxz, maxid, cz[k * 256], ci[k * 256])) { maxz = cz[k * 256]; maxid = ci[k * 256]; maxpos = k; }
      }
    }
    const int m = min(nhit, K);
#pragma unroll
    for (int k = 0; k < K; ++k)
      if (k < m) { L.z[k] = cz[k * 256]; L.id[k] = ci[k * 256]; L.q[k] = cq[k * 256]; }
    L.cnt = m;
  }
  L.sort();
  const int size = L.cnt;
  float zmax = -1000.0f;
#pragma unroll
  for (int k = 0; k < K; ++k)
    if (k < size) zmax = L.z[k];
  const float z0 = L.z[0];
  bool alive = true;
  int oi[K];
  float oz[K], oq[K];
#pragma unroll
  for (int k = 0; k < K; ++k) {
    alive = alive && (k < size) && !(__fsub_rn(L.z[k], z0) > depth_merging_thres);
    oi[k] = alive ? L.id[k] : -1;
    oz[k] = alive ? L.z[k] : -1.0f;
    oq[k] = alive ? L.q[k] : -1.0f;
  }
  const size_t pix = ((size_t)n * S + row) * S + col;
  out_occ[pix] = (size > 0 && (occ_inclusive ? (zmax >= 0.0f) : (zmax > 0.0f))) ? 1.0f : 0.0f;
  int* pi = out_idx + pix * K;
  float* pz = out_z + pix * K;
  float* pq = out_q + pix * K;
  if constexpr (K % 4 == 0) {
#pragma unroll
    for (int k = 0; k < K; k += 4) {
      *reinterpret_cast<int4*>(pi + k) = make_int4(oi[k], oi[k + 1], oi[k + 2], oi[k + 3]);
      *reinterpret_cast<float4*>(pz + k) = make_float4(oz[k], oz[k + 1], oz[k + 2], oz[k + 3]);
      *reinterpret_cast<float4*>(pq + k) = make_float4(oq[k], oq[k + 1], oq[k + 2], oq[k + 3]);
    }
  } else if constexpr (K % 2 == 0) {
#pragma unroll
    for (int k = 0; k < K; k += 2) {
      *reinterpret_cast<int2*>(pi + k) = make_int2(oi[k], oi[k + 1]);
      *reinterpret_cast<float2*>(pz + k) = make_float2(oz[k], oz[k + 1]);
      *reinterpret_cast<float2*>(pq + k) = make_float2(oq[k], oq[k + 1]);
    }
  } else {
#pragma unroll
    for (int k = 0; k < K; ++k) { pi[k] = oi[k]; pz[k] = oz[k]; pq[k] = oq[k]; }
  }
  if (epi.visible && oi[0] >= 0) {
#pragma unroll
    for (int k = 0; k < K; ++k)
      if (oi[k] >= 0) epi.visible[oi[k]] = 1;
  }
  if (epi.img) {
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    float sw = 0.f;
#pragma unroll
    for (int k = 0; k < K; ++k) {
      float wgt = 0.f;
      if (oi[k] >= 0) {
        wgt = expf(-0.5f * oq[k]) * (epi.scaler ? epi.scaler[oi[k]] : 1.0f);
        sw += wgt;
        const float* f = epi.feat + (size_t)oi[k] * epi.feat_stride;
#pragma unroll
        for (int c = 0; c < 4; ++c)
          if (c < epi.C) acc[c] += wgt * f[c];
      }
      if (epi.weights) epi.weights[pix * K + k] = wgt;
    }
    const float den = fmaxf(sw, epi.eps);
    float* o = epi.img + pix * (epi.C + 1);
#pragma unroll
    for (int c = 0; c < 4; ++c)
      if (c < epi.C) o[c] = acc[c] / den;
    o[epi.C] = (size > 0 && (occ_inclusive ? (zmax >= 0.0f) : (zmax > 0.0f))) ? 1.0f : 0.0f;
  }
}

// Generic K (17..150, rasterization_utils.cuh:18): same algorithm, list in local memory.
__global__ void __launch_bounds__(256)
splat_raster_bigk_kernel(const float4* __restrict__ recs, const int* __restrict__ tile_off,
                         const int* __restrict__ tile_cnt, int S, int T, int K,
                         float depth_merging_thres, int occ_inclusive, int* __restrict__ out_idx,
                         float* __restrict__ out_z, float* __restrict__ out_q,
                         float* __restrict__ out_occ) {
  constexpr int KMAX = 150;
  const int tile = blockIdx.x;
  const int n = tile / (T * T);
  const int tr = tile - n * T * T;
  const int ty = tr / T, tx = tr - ty * T;
  const int col = tx * TILE + (threadIdx.x & 15), row = ty * TILE + (threadIdx.x >> 4);
  if (col >= S || row >= S) return;
  const float fS = (float)S;
  const float xf = pix_to_ndc(S - 1 - col, fS), yf = pix_to_ndc(S - 1 - row, fS);
  float lz[KMAX], lq[KMAX];
  int li[KMAX];
  int size = 0;
  const int nrec = tile_cnt[tile];
  const float4* rb = recs + (long long)tile_off[tile] * REC_F4;
  for (int s = 0; s < nrec; ++s) {
    const float4 a0 = rb[s * REC_F4 + 0], a1 = rb[s * REC_F4 + 1], a2 = rb[s * REC_F4 + 2];
    const float dx = __fsub_rn(xf, a0.x), dy = __fsub_rn(yf, a0.y);
    if (fabsf(dx) > a1.w || fabsf(dy) > a2.x) continue;
    const float q = __fmaf_rn(__fmul_rn(a1.y, dy), dy,
                              __fmaf_rn(__fmul_rn(a0.w, dx), dx, __fmul_rn(__fmul_rn(a1.x, dx), dy)));
    if (q > a1.z) continue;
    const int cid = __float_as_int(a2.y);
    if (size == K && !zid_less(a0.z, cid, lz[K - 1], li[K - 1])) continue;
    int k = size < K ? size : K - 1;
    while (k > 0 && zid_less(a0.z, cid, lz[k - 1], li[k - 1])) {
      lz[k] = lz[k - 1]; li[k] = li[k - 1]; lq[k] = lq[k - 1];
      --k;
    }
    lz[k] = a0.z; li[k] = cid; lq[k] = q;
    if (size < K) ++size;
  }
  const size_t pix = ((size_t)n * S + row) * S + col;
  const float zmax = size > 0 ? lz[size - 1] : -1000.0f;
  out_occ[pix] = (size > 0 && (occ_inclusive ? (zmax >= 0.0f) : (zmax > 0.0f))) ? 1.0f : 0.0f;
  bool alive = true;
  for (int k = 0; k < K; ++k) {
    alive = alive && (k < size) && !(__fsub_rn(lz[k], lz[0]) > depth_merging_thres);
    out_idx[pix * K + k] = alive ? li[k] : -1;
    out_z[pix * K + k] = alive ? lz[k] : -1.0f;
    out_q[pix * K + k] = alive ? lq[k] : -1.0f;
  }
}

// points_per_bin of RasterizePointsCoarseCudaKernel (rasterize_points.cu:353-412), which the
// reference computes and drops: bins of bin_size pixels in NDC index space (NOT flipped), extents
// PixToNdc(b*bin) - 1/S .. PixToNdc((b+1)*bin - 1) + 1/S in fp32, inclusive overlap, z >= 0.
__global__ void __launch_bounds__(256)
splat_bin_count_kernel(const float* __restrict__ points, const float* __restrict__ radii,
                       const int64_t* __restrict__ first_idx, const int64_t* __restrict__ num_points,
                       int S, int bin_size, int B, int* __restrict__ bin_cnt) {
  const int n = blockIdx.y;
  const long long first = first_idx[n], num = num_points[n];
  const float fS = (float)S;
  const float half = __fdiv_rn(1.0f, fS);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < num;
       i += (long long)gridDim.x * blockDim.x) {
    const long long p = first + i;
    const float px = points[3 * p], py = points[3 * p + 1], pz = points[3 * p + 2];
    if (pz < 0) continue;
    const float px0 = __fsub_rn(px, radii[2 * p]), px1 = __fadd_rn(px, radii[2 * p]);
    const float py0 = __fsub_rn(py, radii[2 * p + 1]), py1 = __fadd_rn(py, radii[2 * p + 1]);
    for (int by = 0; by < B; ++by) {
      const float by0 = __fsub_rn(pix_to_ndc(by * bin_size, fS), half);
      const float by1 = __fadd_rn(pix_to_ndc((by + 1) * bin_size - 1, fS), half);
      if (!((py0 <= by1) && (by0 <= py1))) continue;
      for (int bx = 0; bx < B; ++bx) {
        const float bx0 = __fsub_rn(pix_to_ndc(bx * bin_size, fS), half);
        const float bx1 = __fadd_rn(pix_to_ndc((bx + 1) * bin_size - 1, fS), half);
        if ((px0 <= bx1) && (bx0 <= px1)) atomicAdd(&bin_cnt[((size_t)n * B + by) * B + bx], 1);
      }
    }
  }
}

// number of (pixel, point) pairs that pass CheckPixelInsidePoint -- the "pixel-splat" unit of the
// throughput metric.  One thread per point sweeps the point's own pixel window.
__global__ void __launch_bounds__(256)
splat_pair_count_kernel(const float* __restrict__ points, const float* __restrict__ ellipse,
                        const float* __restrict__ cutoff, const float* __restrict__ radii, long long P,
                        int S, unsigned long long* __restrict__ total) {
  const float fS = (float)S;
  unsigned long long mine = 0;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < P;
       p += (long long)gridDim.x * blockDim.x) {
    const float px = points[3 * p], py = points[3 * p + 1], pz = points[3 * p + 2];
    if (!(pz >= 0.0f)) continue;
    const float rx = radii[2 * p], ry = radii[2 * p + 1];
    const float a = ellipse[3 * p], b = ellipse[3 * p + 1], c = ellipse[3 * p + 2], cut = cutoff[p];
    int xl, xh, yl, yh;
    pixel_range(px, rx, S, fS, xl, xh);
    pixel_range(py, ry, S, fS, yl, yh);
    for (int yi = yl; yi <= yh; ++yi) {
      const float dy = __fsub_rn(pix_to_ndc(yi, fS), py);
      if (fabsf(dy) > ry) continue;
      for (int xi = xl; xi <= xh; ++xi) {
        const float dx = __fsub_rn(pix_to_ndc(xi, fS), px);
        if (fabsf(dx) > rx) continue;
        const float q = __fmaf_rn(__fmul_rn(c, dy), dy, __fmaf_rn(__fmul_rn(a, dx), dx, __fmul_rn(__fmul_rn(b, dx), dy)));
        if (!(q > cut)) ++mine;
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
  if ((threadIdx.x & 31) == 0 && mine) atomicAdd(total, mine);
}

template <int K>
static void launch_raster(int variant, int tiles, cudaStream_t st, const float4* recs, const int* off,
                          const int* cnt, int S, int T, float thres, int occ_incl, int* oi, float* oz,
                          float* oq, float* oo, const RasterEpi& epi) {
  if (variant != 1) {
    // v2.  K <= 8: 20 slots per pixel column (60 KB) + ONE 12 KB staging buffer = 73 KB per CTA, three CTAs per SM --
    // the other CTAs of the SM cover the record copy that a second buffer would overlap (measured against 24 slots
    // + two buffers = 97 KB, two CTAs per SM: 0.486 -> 0.468 ms for the fused forward of BASELINE config 4; that
    // form stays selectable as variant 2).  Larger K: K + 16 slots, two buffers.
    if (K <= 8 && variant != 2) {
      constexpr int C = 20;
      const int smem2 = 3 * C * 256 * (int)sizeof(float);
      cudaFuncSetAttribute(splat_raster_v2_kernel<K, C, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem2);
      splat_raster_v2_kernel<K, C, 1><<<tiles, 256, smem2, st>>>(recs, off, cnt, S, T, thres, occ_incl, oi, oz, oq, oo, epi);
      return;
    }
    constexpr int C = (K <= 8) ? 24 : K + 16;
    const int smem2 = 3 * C * 256 * (int)sizeof(float);
    cudaFuncSetAttribute(splat_raster_v2_kernel<K, C, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem2);
    splat_raster_v2_kernel<K, C, STAGES><<<tiles, 256, smem2, st>>>(recs, off, cnt, S, T, thres, occ_incl, oi, oz, oq, oo, epi);
    return;
  }
  const int smem = 3 * K * 256 * (int)sizeof(float);
  // static + dynamic shared memory passes 48 KB from K = 12: opt in (per device, so on every launch)
  cudaFuncSetAttribute(splat_raster_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  splat_raster_kernel<K><<<tiles, 256, smem, st>>>(recs, off, cnt, S, T, thres, occ_incl, oi, oz, oq, oo);
}

struct SplatWs {
  int* tile_cnt;
  int* tile_off;
  int* tile_cur;
  int* ticket;
  int* total;
  void* scan_ws;
  size_t scan_bytes;
  size_t bytes;
};

static SplatWs carve_splat_ws(void* ws, int N, int S) {
  const int T = div_up(S, TILE);
  const size_t nt = (size_t)N * T * T;
  SplatWs w;
  size_t off = 0;
  auto take = [&](size_t b) { void* p = ws ? (char*)ws + off : nullptr; off += align_up(b); return p; };
  w.tile_cnt = (int*)take(nt * 4);
  w.tile_cur = (int*)take(nt * 4);     // adjacent to tile_cnt: one memset clears both
  w.ticket = (int*)take(4);            // ... and the finish ticket of the count kernel
  w.tile_off = (int*)take(nt * 4);
  w.total = (int*)take(4);
  w.scan_bytes = scan_ws_bytes((int)nt, 1);
  w.scan_ws = take(w.scan_bytes);
  w.bytes = off;
  return w;
}

}  // namespace isob200

using namespace isob200;

extern "C" {

size_t isob200_splat_ws_bytes(int N, int S) { return carve_splat_ws(nullptr, N, S).bytes; }
int isob200_splat_record_bytes(void) { return REC_F4 * 16; }

// Phase 1 of DSS._C.splat_points: per-tile record counts + offsets.  total_out (device int)
// receives the number of (tile, point) records; the caller reads it back once to size `recs`.
int isob200_splat_bin(const float* points, const float* radii, const int64_t* first_idx,
                      const int64_t* num_points, int N, long long P, long long max_points_per_cloud,
                      int S, void* ws, size_t ws_bytes, int* total_out, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  ISO_CHECK_ARG(N >= 0 && S > 0 && P >= 0, "splat_bin: bad sizes");
  SplatWs w = carve_splat_ws(ws, N, S);
  if (ws == nullptr || ws_bytes < w.bytes) {
    set_error("splat_bin: workspace too small (%zu < %zu)", ws_bytes, w.bytes);
    return ISOB200_ERR_WORKSPACE;
  }
  ISO_CHECK_ARG(total_out, "splat_bin: null total_out");
  const int T = div_up(S, TILE);
  const size_t nt = (size_t)N * T * T;
  ISO_CHECK_ARG(nt < (1u << 31), "splat_bin: too many tiles");
  ISO_CUDA(cudaMemsetAsync(w.tile_cnt, 0, (char*)w.tile_off - (char*)w.tile_cnt, st));
  if (N > 0 && P > 0 && max_points_per_cloud > 0) {
    ISO_CHECK_ARG(points && radii && first_idx && num_points, "splat_bin: null pointer");
    if (T * T <= PRIV_MAX_TILES && P / N >= PRIV_CHUNK) {
      // chunks sized for the AVERAGE view (max_points_per_cloud is only an upper bound, usually P itself): a
      // larger view gets proportionally larger chunks, the grid is the same for every view
      const int pbx = (int)min((long long)div_up(P / N, PRIV_CHUNK), (long long)kNumSMs * 4);
      const bool tail = nt <= (size_t)LAST_CTA_SCAN_MAX;
      splat_tile_count_priv_kernel<<<dim3(pbx, N), 256, (size_t)T * T * sizeof(int), st>>>(
          points, radii, first_idx, num_points, S, T, w.tile_cnt, tail ? w.ticket : nullptr, w.tile_off, w.total,
          total_out);
      ISO_CHECK_LAUNCH("splat_tile_count_priv_kernel");
      if (tail) return ISOB200_OK;
    } else {
      int bx = grid_for(max_points_per_cloud, 256, 8);
      if (N > 1) bx = max(1, min(bx, (kNumSMs * 8 + N - 1) / N));
      splat_tile_count_kernel<<<dim3(bx, N), 256, 0, st>>>(points, radii, first_idx, num_points, S, T,
                                                          w.tile_cnt);
      ISO_CHECK_LAUNCH("splat_tile_count_kernel");
    }
  }
  if (nt > 0) {
    int rc = exclusive_scan_i32(w.tile_cnt, w.tile_off, (int)nt, 1, (long long)nt, (long long)nt, w.scan_ws,
                                w.scan_bytes, st);
    if (rc) return rc;
  }
  splat_total_kernel<<<1, 32, 0, st>>>(w.tile_cnt, w.tile_off, (int)nt, w.total);
  ISO_CHECK_LAUNCH("splat_total_kernel");
  ISO_CUDA(cudaMemcpyAsync(total_out, w.total, sizeof(int), cudaMemcpyDeviceToDevice, st));
  return ISOB200_OK;
}

// Phase 2 of DSS._C.splat_points (rasterize_points.h:461-525): fill the per-tile record lists
// and rasterise.  `ws` must be the workspace isob200_splat_bin just filled; `recs` holds
// `capacity` records of isob200_splat_record_bytes() each (capacity >= the total read back).
// Outputs are fully written: idx int32 (N,S,S,K), zbuf/qvalue f32 (N,S,S,K), occ f32 (N,S,S).
// occ_inclusive: bit 0: 1 = naive-kernel rule q_max_z >= 0 (bin_size == 0), 0 = fine-kernel rule > 0;
//                bits 8..9: raster variant, 0 = v2 (record-centric hit generation, default), 2 = v2 with the
//                round-2a shared-memory budget (24 slots, two staging buffers),
//                1 = v1 (pixel-centric with warp-level culling).  Results are identical.
int isob200_splat_forward_fused(const float* points, const float* ellipse, const float* cutoff,
                                const float* radii, const int64_t* first_idx, const int64_t* num_points,
                                int N, long long P, long long max_points_per_cloud, int S, int K,
                                float depth_merging_thres, int occ_inclusive, void* ws, size_t ws_bytes,
                                void* recs, long long capacity, int* out_idx, float* out_zbuf,
                                float* out_qvalue, float* out_occ, const float* scaler, const float* feat,
                                int feat_stride, int C, float eps, float* out_img, float* out_weights,
                                unsigned char* visible, void* stream_);

int isob200_splat_forward(const float* points, const float* ellipse, const float* cutoff,
                          const float* radii, const int64_t* first_idx, const int64_t* num_points,
                          int N, long long P, long long max_points_per_cloud, int S, int K,
                          float depth_merging_thres, int occ_inclusive, void* ws, size_t ws_bytes,
                          void* recs, long long capacity, int* out_idx, float* out_zbuf,
                          float* out_qvalue, float* out_occ, void* stream_) {
  return isob200_splat_forward_fused(points, ellipse, cutoff, radii, first_idx, num_points, N, P,
                                     max_points_per_cloud, S, K, depth_merging_thres, occ_inclusive, ws, ws_bytes,
                                     recs, capacity, out_idx, out_zbuf, out_qvalue, out_occ, nullptr, nullptr, 0, 0,
                                     0.f, nullptr, nullptr, nullptr, stream_);
}

// isob200_splat_forward with the RGBA blend (isob200_splat_blend: scaler (P) or NULL, feat (P, feat_stride) with
// C <= 4 channels, eps; out_img (N,S,S,C+1), out_weights (N,S,S,K) or NULL) and / or the per-point visibility of
// isob200_splat_visibility(mask = NULL) (visible (P) uint8, zeroed by the caller) computed in the raster kernel's
// epilogue.  feat == NULL skips the blend, visible == NULL the visibility.  Needs the default raster variant and
// K <= 16 when either is requested.  Same bits as the separate kernels.
int isob200_splat_forward_fused(const float* points, const float* ellipse, const float* cutoff,
                                const float* radii, const int64_t* first_idx, const int64_t* num_points,
                                int N, long long P, long long max_points_per_cloud, int S, int K,
                                float depth_merging_thres, int occ_inclusive, void* ws, size_t ws_bytes,
                                void* recs, long long capacity, int* out_idx, float* out_zbuf,
                                float* out_qvalue, float* out_occ, const float* scaler, const float* feat,
                                int feat_stride, int C, float eps, float* out_img, float* out_weights,
                                unsigned char* visible, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  RasterEpi epi = {scaler, feat, feat_stride, C, eps, feat ? out_img : nullptr, feat ? out_weights : nullptr, visible};
  if (feat || visible) {
    ISO_CHECK_ARG(K <= 16 && ((occ_inclusive >> 8) & 3) != 1,
                  "splat_forward_fused: the fused epilogue needs the default raster variant and K <= 16");
    ISO_CHECK_ARG(!feat || (out_img && C >= 1 && C <= 4 && feat_stride >= C), "splat_forward_fused: bad blend arguments");
  }
  ISO_CHECK_ARG(N >= 0 && S > 0 && P >= 0, "splat_forward: bad sizes");
  ISO_CHECK_ARG(K >= 1 && K <= 150, "Must have points_per_pixel <= 150");
  if (N == 0) return ISOB200_OK;
  SplatWs w = carve_splat_ws(ws, N, S);
  if (ws == nullptr || ws_bytes < w.bytes) {
    set_error("splat_forward: workspace too small (%zu < %zu)", ws_bytes, w.bytes);
    return ISOB200_ERR_WORKSPACE;
  }
  ISO_CHECK_ARG(out_idx && out_zbuf && out_qvalue && out_occ, "splat_forward: null output");
  ISO_CHECK_ARG(capacity == 0 || recs, "splat_forward: null record buffer");
  ISO_CHECK_ARG(((uintptr_t)recs & 15) == 0, "splat_forward: record buffer must be 16-byte aligned");
  const int T = div_up(S, TILE);
  const int tiles = N * T * T;
  if (P > 0 && max_points_per_cloud > 0 && capacity > 0) {
    ISO_CHECK_ARG(points && ellipse && cutoff && radii && first_idx && num_points, "splat_forward: null pointer");
    int bx = grid_for(max_points_per_cloud, 256, 8);
    if (N > 1) bx = max(1, min(bx, (kNumSMs * 8 + N - 1) / N));
    splat_tile_fill_kernel<<<dim3(bx, N), 256, 0, st>>>(points, ellipse, cutoff, radii, first_idx, num_points,
                                                       S, T, w.tile_off, w.tile_cur, capacity, (float4*)recs);
    ISO_CHECK_LAUNCH("splat_tile_fill_kernel");
  }
  const int variant = (occ_inclusive >> 8) & 3;
  occ_inclusive &= 1;
  const float4* r = (const float4*)recs;
  int* oi = out_idx; float* oz = out_zbuf; float* oq = out_qvalue; float* oo = out_occ;
  const float th = depth_merging_thres;
#define RK(KK) case KK: launch_raster<KK>(variant, tiles, st, r, w.tile_off, w.tile_cnt, S, T, th, occ_inclusive, oi, oz, oq, oo, epi); break;
  switch (K) {
    RK(1) RK(2) RK(3) RK(4) RK(5) RK(6) RK(7) RK(8) RK(9) RK(10) RK(11) RK(12) RK(13) RK(14) RK(15) RK(16)
    default:
      splat_raster_bigk_kernel<<<tiles, 256, 0, st>>>(r, w.tile_off, w.tile_cnt, S, T, K, th, occ_inclusive, oi,
                                                     oz, oq, oo);
  }
#undef RK
  ISO_CHECK_LAUNCH("splat_raster_kernel");
  return ISOB200_OK;
}

// total_out (device uint64, zeroed here) = number of pixel-splats of the packed splat set.
int isob200_splat_count_pairs(const float* points, const float* ellipse, const float* cutoff,
                              const float* radii, long long P, int S, unsigned long long* total_out,
                              void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  ISO_CHECK_ARG(total_out && S > 0 && P >= 0, "splat_count_pairs: bad argument");
  ISO_CUDA(cudaMemsetAsync(total_out, 0, sizeof(unsigned long long), st));
  if (P == 0) return ISOB200_OK;
  ISO_CHECK_ARG(points && ellipse && cutoff && radii, "splat_count_pairs: null pointer");
  splat_pair_count_kernel<<<grid_for(P, 256, 8), 256, 0, st>>>(points, ellipse, cutoff, radii, P, S, total_out);
  ISO_CHECK_LAUNCH("splat_pair_count_kernel");
  return ISOB200_OK;
}

// (N,B,B) int32 points_per_bin of the reference's coarse pass, B = 1 + (S-1)/bin_size,
// indexed [n][by][bx] in NDC bin order like the reference's bin_points (rasterize_points.cu:402-412).
int isob200_splat_bin_counts(const float* points, const float* radii, const int64_t* first_idx,
                             const int64_t* num_points, int N, long long max_points_per_cloud, int S,
                             int bin_size, int* bin_cnt, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  ISO_CHECK_ARG(N >= 0 && S > 0 && bin_size > 0, "splat_bin_counts: bad sizes");
  const int B = 1 + (S - 1) / bin_size;
  if (N == 0) return ISOB200_OK;
  ISO_CHECK_ARG(bin_cnt, "splat_bin_counts: null output");
  ISO_CUDA(cudaMemsetAsync(bin_cnt, 0, (size_t)N * B * B * sizeof(int), st));
  if (max_points_per_cloud <= 0) return ISOB200_OK;
  int bx = grid_for(max_points_per_cloud, 256, 8);
  if (N > 1) bx = max(1, min(bx, (kNumSMs * 8 + N - 1) / N));
  splat_bin_count_kernel<<<dim3(bx, N), 256, 0, st>>>(points, radii, first_idx, num_points, S, bin_size, B,
                                                     bin_cnt);
  ISO_CHECK_LAUNCH("splat_bin_count_kernel");
  return ISOB200_OK;
}

}  // extern "C"
