/* isob200.h -- C ABI of libisob200.so, the B200 (sm_100a) implementation of the iso-points hot path.
 *
 * Every entry point
 *   - takes raw DEVICE pointers + sizes + a cudaStream_t (passed as void*), no torch types;
 *   - never allocates: outputs and workspaces are caller-owned (a *_ws_bytes() query precedes
 *     each call that needs scratch);
 *   - is stream-ordered on the given stream and returns an int status: 0 = ok, 1 = invalid
 *     argument, 2 = CUDA error, 3 = workspace too small; isob200_last_error() returns the
 *     message of the last failure on the calling thread (the Python layer raises RuntimeError
 *     with it, as the reference's TORCH_CHECK / AT_CUDA_CHECK would).
 * Each declaration cites the reference interface it replaces (paths relative to the
 * yifita/iso-points tree).  The reference-side bindings are shown in INTEGRATION.md.
 */
#ifndef ISOB200_H_
#define ISOB200_H_
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- library / error channel -------------------------------------------------------- */
const char* isob200_last_error(void);
int isob200_abi_version(void);
int isob200_compiled_arch(void); /* 1000 = sm_100a */
long long isob200_launch_count(void); /* kernels launched by this library so far (bench.py gpu_launches) */

/* ---- exclusive scan: prefix_sum.prefix_sum_cuda(cnt, num_cells, off)
 *      external/FRNN/external/prefix_sum/prefix_sum.cu:74-87 (prefix_sum.h:18-20);
 *      batched over `rows`, on the caller's stream, no cudaMalloc --------------------- */
size_t isob200_exclusive_scan_ws_bytes(int n, int rows);
int isob200_exclusive_scan_i32(const int* in, int* out, int n, int rows, long long in_stride,
                               long long out_stride, void* ws, size_t ws_bytes, void* stream);

/* ---- FRNN grid: external/FRNN/frnn/frnn.py:55-71 (grid params host loop),
 *      frnn._C.insert_points_cuda  (csrc/grid/grid.cu:135-184, ext.cpp:10),
 *      frnn._C.counting_sort_cuda  (csrc/grid/counting_sort.cu:73-125),
 *      and a fused deterministic build replacing insert + prefix_sum + counting_sort -- */
int isob200_frnn_grid_params(const float* points, const int64_t* lengths, const float* rs, int N,
                             int P, int D, double radius_cell_ratio, float* params, int* g_max,
                             void* ws, size_t ws_bytes, void* stream);
/* the same for a caller that sizes its (N, g_cap) cell table before reading g_max back: a cloud that needs more
 * cells gets a one-cell grid no query reaches (everything stays in bounds); g_max still reports the true size */
int isob200_frnn_grid_params_capped(const float* points, const int64_t* lengths, const float* rs, int N,
                                    int P, int D, double radius_cell_ratio, int g_cap, float* params, int* g_max,
                                    void* ws, size_t ws_bytes, void* stream);
/* min / max over the live rows of each cloud, out (N,2,D), lengths on the device: the bounding box of
 * DSS/models/levelset_sampling.py:254 (`points.view(-1,3).max(0) - min(0)`) without a host-side survivor count */
int isob200_points_bbox(const float* points, const int64_t* lengths, int N, int P, int D, float* out, void* ws,
                        size_t ws_bytes, void* stream);
int isob200_frnn_insert_points(const float* points, const int64_t* lengths, const float* params,
                               int* grid_cnt, int* grid_cell, int* grid_idx, int N, int P, int D,
                               int G, void* stream);
int isob200_frnn_counting_sort(const float* points, const int64_t* lengths, const int* grid_cell,
                               const int* grid_idx, const int* grid_off, float* sorted_points,
                               int* sorted_idxs, int N, int P, int D, int G, void* stream);
size_t isob200_frnn_build_ws_bytes(int N, int P, int G);
int isob200_frnn_build(const float* points, const int64_t* lengths, const float* params, int N, int P,
                       int D, int G, int* cell_off, float* sorted_points, int* sorted_idxs, void* ws,
                       size_t ws_bytes, void* stream);

/* ---- FRNN query: frnn._C.find_nbrs_cuda (csrc/grid/grid.cu:384-440),
 *      frnn.frnn_gather (frnn.py:304-352) and its autograd, frnn._C.frnn_backward_cuda
 *      (csrc/backward/backward.cu:76-147) ------------------------------------------------ */
/* group_width: lanes cooperating on one query (0 = auto: smallest of 8/16/32 >= K), optionally OR-ed with a
 * traversal mode << 8: 0 auto, 1 exhaustive block scan, 2 pruned best-first (identical results) */
int isob200_frnn_find_nbrs(const float* q_points, const int* q_order, const int64_t* lengths1,
                           const int64_t* lengths2, const float* sorted_points2, const int* cell_off2,
                           const int* sorted_idxs2, const float* params, const float* rs, int N,
                           int P1, int P2, int D, int G, int K, float* dists, void* idxs,
                           int idx_is_i64, int group_width, void* stream);
int isob200_frnn_gather(const float* x, const void* idxs, int idx_is_i64, int N, int M, int L, int K,
                        int U, float* out, void* stream);
int isob200_frnn_gather_backward(const float* grad_out, const void* idxs, int idx_is_i64, int N, int M,
                                 int L, int K, int U, float* grad_x, void* stream);
int isob200_frnn_backward(const float* points1, const float* points2, const int64_t* lengths1,
                          const int64_t* lengths2, const int64_t* idxs, const float* grad_dists, int N,
                          int P1, int P2, int D, int K, float* grad_points1, float* grad_points2,
                          void* stream);

/* ---- level-set projection: UniformProjection._project_points
 *      (DSS/models/levelset_sampling.py:290-351); one call per Newton iteration ---------- */
size_t isob200_project_step_ws_bytes(int A);
int isob200_project_step(float* points, float* normals, unsigned char* not_converged,
                         const int* act_in, int A, const int* a_dev, const float* sdf, const float* grad,
                         float tol, float max_step, int do_update, int* act_out, float* next_points,
                         int* count_out, void* ws, size_t ws_bytes, void* stream);
/* ---- ray marching: SphereTracing.project_points (DSS/models/levelset_sampling.py:679-808); one call per
 *      iteration, conventions as isob200_project_step: rays advance by alpha * sdf along dirs (step clamped
 *      to max_step), retire when |sdf| <= active_tol or when the step would leave the sphere of radius `bound` */
int isob200_trace_step(float* points, const float* dirs, float* eval, float* grad_out, const int* act_in, int A,
                       const int* a_dev, const float* sdf, const float* grad, float active_tol, float alpha,
                       float max_step, float bound, int do_update, int* act_out, float* next_points,
                       int* count_out, void* ws, size_t ws_bytes, void* stream);
int isob200_gather_rows3(const float* src, const int* idx, int A, float* dst, void* stream);
/* _filter_projection_result (levelset_sampling.py:59-65) for one packed cloud: the converged rows of
 * points / normals, in order, into out_* (>= M rows); *count_out = survivors.  ws as for project_step.
 * normals / out_normals may both be NULL (rows of one array only). */
int isob200_compact_valid(const float* points, const float* normals, const unsigned char* valid,
                          int M, float* out_points, float* out_normals, int* count_out, void* ws,
                          size_t ws_bytes, void* stream);
int isob200_project_sphere(float* points, float* normals, unsigned char* valid, long long M,
                           float radius, float tol, float max_step, int max_iters, void* stream);

/* ---- fused SIREN SDF value + input gradient: UniformProjection._compute_sdf_and_grad
 *      (DSS/models/levelset_sampling.py:142-170) for SDF modules that are the reference's
 *      Siren MLP (DSS/models/common.py:56-165; hidden width 256).  `pack` converts the fp32
 *      parameters (w0 (256,3), b0 (256), w_hidden (L,256,256), b_hidden (L,256), w_last (256),
 *      b_last (1); biases may be NULL) into the tensor-core operand images once per parameter
 *      version; `sdf_grad` evaluates sdf (n) and d sdf / d x (n,3) for x (n,3).  n_dev, when
 *      non-NULL, is a device int with the live row count (<= n_max) so a projection loop can
 *      run without reading the active count back.  dbg (NULL in production) receives the raw
 *      (128,256) accumulator of the first tile's GEMM number dbg_gemm (parity tests). ------- */
size_t isob200_siren_blob_bytes(int n_hidden);
size_t isob200_siren_pack_ws_bytes(void);
size_t isob200_siren_scratch_bytes(int n_hidden);
int isob200_siren_pack(const float* w0, const float* b0, const float* w_hidden, const float* b_hidden,
                       const float* w_last, const float* b_last, float omega0, float omega, int hidden,
                       int n_hidden, void* blob, size_t blob_bytes, void* ws, size_t ws_bytes,
                       void* stream);
int isob200_siren_sdf_grad(const float* x, int n_max, const int* n_dev, const void* blob, int n_hidden,
                           float* sdf, float* grad, void* scratch, size_t scratch_bytes, float* dbg,
                           int dbg_gemm, void* stream);

/* sdf_grad + project_step in one kernel (one Newton iteration, levelset_sampling.py:313-342, for the
 * fused decoder): x = positions of the active rows (points[act_in], compacted), arguments as in
 * isob200_project_step; *count_out must be zero before the call; act_out / next_points are appended
 * per 128-row tile in completion order (rows are independent: results identical). */
int isob200_siren_project_step(const float* x, int n_max, const int* n_dev, const void* blob, int n_hidden,
                               void* scratch, size_t scratch_bytes, float* points, float* normals,
                               unsigned char* not_converged, const int* act_in, float tol, float max_step,
                               int do_update, int* act_out, float* next_points, int* count_out,
                               void* stream);

/* SDF value only (forward half of isob200_siren_sdf_grad: half the tensor work, no tape); same bits as its sdf */
int isob200_siren_sdf(const float* x, int n_max, const int* n_dev, const void* blob, int n_hidden, float* sdf,
                      void* scratch, size_t scratch_bytes, void* stream);
/* isob200_trace_step with the SDF evaluation of the reference's Siren decoder fused in (forward half of
 * the network only: the march needs no gradient); conventions as isob200_siren_project_step */
int isob200_siren_trace_step(const float* x, int n_max, const int* n_dev, const void* blob, int n_hidden,
                             void* scratch, size_t scratch_bytes, float* points, const float* dirs, float* eval,
                             const int* act_in, float active_tol, float alpha, float max_step, float bound,
                             int do_update, int* act_out, float* next_points, int* count_out, void* stream);

/* ---- in-surface sampler: closest point to every ray, Model.sample_offsurface_using_isopoints
 *      (DSS/models/combined_modeling.py:325-352: the two (R,M) dist_to_ray matrices + topk(k = 1)).  For ray r
 *      with origin o (n_origins = 1: shared by all rays, else R) and unit direction d, over points p (M,3):
 *      t = (p - o).d, dist = |p - o|^2 - t^2; outputs at the point of smallest dist (ties: lowest index):
 *      t_sq (R) = t^2, dist (R), idx (R) (-1 and t_sq 0 when M = 0); any output may be NULL ------------- */
int isob200_ray_nearest_point(const float* origins, int n_origins, const float* dirs, int R, const float* points,
                              int M, float* t_sq, float* dist, int* idx, void* stream);

/* tuning knob: persistent CTAs per SIREN launch (1..148, default 148 = one per SM); returns the old value */
int isob200_siren_set_max_ctas(int n);
/* tuning knob: the CTAs on odd SMs of a value + gradient launch with >= min_tiles 128-row tiles start `cycles` SM
 * cycles late, so that half of the chip writes reverse-mode tape while the other half consumes it and the live
 * tape stays L2 resident (cycles < 0: half a tile period, the default; 0: off; min_tiles <= 0: keep); returns the
 * old cycle setting */
int isob200_siren_set_stagger(int cycles, int min_tiles);

/* ---- uniform resampling: UniformProjection.resample, one sample_iter
 *      (DSS/models/levelset_sampling.py:259, 268-284) ------------------------------------ */
int isob200_resample_step(const float* q_points, const float* points, const float* normals,
                          const void* idxs, int idx_is_i64, int idx_stride, int k_offset,
                          const float* inv_sigma, int N, int Pq, int P, int K, float* out, void* stream);
int isob200_normalize_rows3(const float* x, long long M, float eps, float* out, void* stream);

/* ---- DSS elliptical splat rasteriser, forward: DSS._C.splat_points
 *      (DSS/csrc/rasterize_points.h:461-525 -> RasterizePoints{Naive,Coarse,Fine}Cuda,
 *      rasterize_points.cu:214-285, 434-500, 599-667).  Two phases because the per-tile record
 *      buffer is sized from a device-computed total (one 4-byte read-back by the caller). ---- */
size_t isob200_splat_ws_bytes(int N, int S);
int isob200_splat_record_bytes(void);
int isob200_splat_bin(const float* points, const float* radii, const int64_t* first_idx,
                      const int64_t* num_points, int N, long long P, long long max_points_per_cloud,
                      int S, void* ws, size_t ws_bytes, int* total_out, void* stream);
int isob200_splat_forward(const float* points, const float* ellipse, const float* cutoff,
                          const float* radii, const int64_t* first_idx, const int64_t* num_points,
                          int N, long long P, long long max_points_per_cloud, int S, int K,
                          float depth_merging_thres, int occ_inclusive, void* ws, size_t ws_bytes,
                          void* recs, long long capacity, int* out_idx, float* out_zbuf,
                          float* out_qvalue, float* out_occ, void* stream);
/* isob200_splat_forward with, computed in the raster kernel's epilogue (a pixel's K entries are still in
 * registers there): the RGBA blend of DSS/core/renderer.py:53-78 (arguments of isob200_splat_blend; feat == NULL:
 * no blend) and / or the per-point visibility of DSS/core/rasterizer.py:851-857 (isob200_splat_visibility with
 * mask == NULL; visible (P) uint8 zeroed by the caller, NULL: none).  Default raster variant, K <= 16.  Same bits
 * as the separate entry points. */
int isob200_splat_forward_fused(const float* points, const float* ellipse, const float* cutoff,
                                const float* radii, const int64_t* first_idx, const int64_t* num_points,
                                int N, long long P, long long max_points_per_cloud, int S, int K,
                                float depth_merging_thres, int occ_inclusive, void* ws, size_t ws_bytes,
                                void* recs, long long capacity, int* out_idx, float* out_zbuf,
                                float* out_qvalue, float* out_occ, const float* scaler, const float* feat,
                                int feat_stride, int C, float eps, float* out_img, float* out_weights,
                                unsigned char* visible, void* stream);
/* points_per_bin of the reference's coarse pass (rasterize_points.cu:353-412; computed there,
 * never returned) -- for the "per-tile point counts bit-exact" parity check */
int isob200_splat_bin_counts(const float* points, const float* radii, const int64_t* first_idx,
                             const int64_t* num_points, int N, long long max_points_per_cloud, int S,
                             int bin_size, int* bin_cnt, void* stream);

/* number of (pixel, point) pairs passing CheckPixelInsidePoint (rasterize_points.cu:64-98): the
 * "pixel-splat" unit of the throughput metric */
int isob200_splat_count_pairs(const float* points, const float* ellipse, const float* cutoff,
                              const float* radii, long long P, int S, unsigned long long* total_out,
                              void* stream);

/* ---- splat backward: DSS._C._splat_points_occ_fast_cuda_backward (rasterize_points.h:327-336,
 *      rasterize_points_backward.cu:227-322; mode 0), DSS._C._splat_points_occ_backward
 *      (rasterize_points.h:341-386; mode 1), DSS._C._backward_zbuf (rasterize_points.h:388-419),
 *      per-point visibility (DSS/utils/__init__.py:378-399; DSS/core/rasterizer.py:851-857) ---- */
size_t isob200_splat_occ_backward_ws_bytes(int N, int H, int W, long long total_points);
int isob200_splat_occ_backward(const float* points, const float* radii, const unsigned char* visible,
                               const int64_t* first_idx, const int64_t* num_points, const float* rs,
                               float radii_s, const float* grad_occ, int N, int H, int W,
                               long long total_points, int mode, float* grad_out, int out_stride,
                               void* ws, size_t ws_bytes, void* stream);
size_t isob200_splat_search_radius_ws_bytes(int N);
/* per-view median(visible radii) * radii_s: DSS/core/rasterizer.py:881-884 */
int isob200_splat_search_radius(const float* radii, const unsigned char* visible, const int64_t* first_idx,
                                const int64_t* num_points, int N, long long max_points_per_cloud, float radii_s,
                                float* rs, void* ws, size_t ws_bytes, void* stream);
int isob200_splat_zbuf_backward(const int* idx, const float* grad_zbuf, int N, int H, int W, int K,
                                float* z_grad, int stride, void* stream);
int isob200_splat_visibility(const int* idx, const float* mask, long long npix, int K, long long P,
                             unsigned char* visible, void* stream);

/* ---- RGBA blend: SurfaceSplattingRenderer.forward (DSS/core/renderer.py:53-78: exp(-Q/2)*scaler
 *      weights + pytorch3d NormWeightedCompositor + occupancy as alpha) and its feature gradient */
int isob200_splat_blend(const int* idx, const float* qvalue, const float* occ, const float* scaler,
                        const float* feat, int feat_stride, long long npix, int K, int C, float eps,
                        float* out, float* weights_out, void* stream);
int isob200_splat_blend_backward(const int* idx, const float* weights, const float* grad_out,
                                 long long npix, int K, int C, float eps, float* grad_feat,
                                 int feat_stride, void* stream);

/* ---- per-point EWA splat parameters + renderable filter: the op chains of SurfaceSplatting that
 *      run right before the splat kernel (DSS/core/rasterizer.py).
 *      Packed points (P,3) / normals (P,3), view-major; first_idx / num_points (n_views) int64;
 *      camera matrices row-major in the row-vector convention of pytorch3d (p_hom @ M), either one
 *      per view or a single broadcast one (gather_batch_to_packed, DSS/utils/__init__.py:218-250). */
/* _compute_isotropic_Vrk (:358-386): sq_dists (n_views, P1, K) from the K = 7 self query
 * (slot 0 = the point itself) -> h (P,) = clamp(0.5 max_k>=1, 5e-5, 0.01); clouds with < K points: 1e-3 */
int isob200_ewa_vrk_h(const float* sq_dists, const int64_t* first_idx, const int64_t* num_points, int n_views,
                      long long P1, int K, long long P, float* h, void* stream);
/* _get_per_point_info (:514-563) = _compute_WJk (:438-487) + _compute_variance_and_detMk (:402-436)
 * + _get_ellipse_axis_aligned_radius (:489-512), isotropic V_k^r (:388-400).
 * proj (proj_views,4,4) = cameras.get_full_projection_transform().get_matrix();
 * pixel_var = antialiasing_sigma * (2 / image_size)^2; outputs radii (P,2), ellipse (P,3) =
 * (a, b, c) of a x^2 + b xy + c y^2, cutoff_out (P,), scaler (P,). */
int isob200_ewa_point_params(const float* points, const float* normals, const int64_t* first_idx, int n_views,
                             long long P, const float* proj, int proj_views, const float* vrk_h,
                             float pixel_var, float cutoff, float* radii, float* ellipse, float* cutoff_out,
                             float* scaler, void* stream);
/* filter_renderable (:220-255): _filter_points_with_invalid_depth (:163-218) and, with nmat != NULL,
 * _filter_backface_points (:124-161) as one mask: znear <= z_view <= zfar [and n_view.z < 0].
 * w2v (cam_views,4,4) world-to-view; nmat (cam_views,3,3) the matrix normals are right-multiplied by.
 * mask (P,) uint8; kept (n_views,) int32 = survivors per view (zeroed here). */
int isob200_renderable_mask(const float* points, const float* normals, const int64_t* first_idx, int n_views,
                            long long P, const float* w2v, const float* nmat, int cam_views, float znear,
                            float zfar, unsigned char* mask, int* kept, void* stream);

/* ---- point-set operators around the projection: DSS/utils/point_processing.py
 *      wlop (:35-122: density_P :77-80, one iteration :90-118), upsample (:321-340, one round),
 *      farthest_sampling (:473-499 -> torch_cluster.fps, third party) ------------------------ */
int isob200_wlop_density(const float* pts, const int64_t* idx, int idx_stride, int k_offset,
                         const float* sigma_inv, int N, int P, int K, float* density, void* stream);
int isob200_wlop_step(const float* X, const float* Pc, const int64_t* idx_xp, const int64_t* idx_xx,
                      int xx_stride, int xx_k_offset, const float* density_P, const float* sigma_inv, float mu,
                      int N, int PX, int PP, int K, float* out, void* stream);
/* normals != NULL selects the edge-aware score of EdgeAwareProjection.upsample
 * (DSS/models/levelset_sampling.py:614-628) */
int isob200_upsample_sparsity(const float* pts, const float* normals, float edge_sensitivity,
                              const int64_t* idx, int idx_stride, int k_offset, const int64_t* lengths, int N,
                              int P, int K, float* sparsity, float* child, void* stream);
/* farthest point sampling (torch_cluster.fps behind DSS/utils/point_processing.py:473-499; start point start[n] or
 * 0): mind = (N,P) float scratch, single CTA per cloud.  isob200_fps_ws: same with a scratch of
 * isob200_fps_ws_floats(N,P) floats, which lets one cloud run on all SMs (cooperative launch, slices of the cloud
 * resident in shared memory, one grid barrier per sample); identical indices */
int isob200_fps(const float* pts, const int64_t* lengths, const int64_t* samples, const int64_t* start, int N,
                int P, int Mmax, float* mind, int64_t* out_idx, void* stream);
size_t isob200_fps_ws_floats(int N, int P);
int isob200_fps_ws(const float* pts, const int64_t* lengths, const int64_t* samples, const int64_t* start, int N,
                   int P, int Mmax, float* ws, size_t ws_floats, int64_t* out_idx, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ISOB200_H_ */
