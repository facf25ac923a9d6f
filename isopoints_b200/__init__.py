"""isopoints_b200 -- the iso-points hot path of yifita/iso-points on hand-written sm_100a kernels.

Submodules mirror the reference's operator surface (import them explicitly; nothing is loaded eagerly and the
CUDA library ``libisob200.so`` is opened on first use -- there is no CPU / PyTorch fallback):

    frnn               frnn_grid_points, frnn_gather, _C, prefix_sum_cuda         (external/FRNN, prefix_sum)
    levelset_sampling  UniformProjection, EdgeAwareProjection, SphereTracing, sample_uniform_iso_points
    point_processing   wlop, upsample, resample_uniformly, farthest_sampling
    splat              _C.splat_points & co, rasterize_elliptical_points, blend_rgba (DSS/csrc, core/rasterizer.py)
    ewa                SurfaceSplatting, SurfaceSplattingRenderer, get_visible_points   (DSS/core/rasterizer.py)
    cloud              PointCloudsFilters                                               (DSS/core/cloud.py)
    ray_tracing        RayTracing                                                       (levelset_sampling.py:810-1167)
    offsurface         sample_offsurface_using_isopoints, get_visible_iso_points        (models/combined_modeling.py)
    siren              fused SDF value / input gradient of the reference's Siren decoder (models/common.py)
    dist               point / view sharding over torch.distributed
    structures         pytorch3d-less Pointclouds stand-in and packed / padded helpers
"""
__version__ = "0.1.0"
