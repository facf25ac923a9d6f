"""GPU parity at the BENCHED configuration (BASELINE configs[1], "C2"): 200 000 points, the SDF SURVEY 8d pins
-- the reference's own ``Siren(dim=3, c_dim=0, hidden_size=256, n_layers=7, ...)`` under ``torch.manual_seed(0)``
(tests/helpers.pinned_siren; state_dict equality with the reference class is asserted on CPU) -- through the
fused tcgen05 path of ``UniformProjection``, against

  * the reference's own class evaluated in float64 (golden, value + gradient),
  * the reference's own ``UniformProjection`` run on CPU on the first 4 096 points of the cloud (golden),
  * the opaque-module path of this package (same weights through autograd, fp32, TF32 off), which the golden
    tests pin to the reference, on all 200 000 points.

Tolerance (north_star): fp32 positions within 1e-4 relative (|a - b| <= 1e-4 * max(|b|, 1) per coordinate, i.e.
rtol = 1e-4 with atol = 1e-4 for coordinates inside the unit box).  A random-init SIREN with omega = 30 per layer
is chaotic under Newton iteration: a row whose |sdf| sits within float noise of the tolerance, or whose
trajectory passes a fold of the field, may end elsewhere on the level set in two fp32-accurate evaluations.
Those rows are COUNTED, printed and bounded; every other row is compared per point."""
import numpy as np
import pytest
import torch

from isopoints_b200 import siren
from isopoints_b200.levelset_sampling import ProjectionResult, UniformProjection
from tests.helpers import pinned_siren

pytestmark = pytest.mark.gpu
DEV = "cuda"
RTOL = 1e-4     # north_star: "fp32 point positions ... within 1e-4 rel"


def _c2_cloud(n=200_000):
    g = torch.Generator().manual_seed(1000)          # bench.py's cloud of rank 0
    return ((torch.rand(1, 200_000, 3, generator=g) - 0.5) * 2)[:, :n].contiguous()


def _within(a, b):
    """per-row: every coordinate within RTOL relative (coordinates live in [-1, 1]: atol = RTOL)."""
    return ((a - b).abs() <= RTOL * b.abs().clamp_min(1.0)).all(dim=-1)


class _NoTF32:
    def __enter__(self):
        self.old = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = False

    def __exit__(self, *a):
        torch.backends.cuda.matmul.allow_tf32 = self.old


def test_fused_kernel_on_the_pinned_network_vs_reference_float64(golden):
    g = golden("pinned_siren")
    model = pinned_siren(0).to(DEV)
    x = torch.as_tensor(g["x"])[0, :512].to(DEV)
    s, gr = siren.sdf_and_grad(model, x)
    es = float((s.double().cpu() - torch.as_tensor(g["sdf64"])).abs().max())
    eg = float((gr.double().cpu() - torch.as_tensor(g["grad64"])).abs().max())
    gmax = float(np.abs(g["grad64"]).max())
    print("pinned SIREN vs the reference class in float64: sdf %.2e abs, grad %.2e abs (%.2e rel)" % (es, eg, eg / gmax))
    assert es < 2e-6 and eg < 3e-5 * gmax


def test_value_and_gradient_at_200k_rows_vs_float64_autograd():
    """All 200 000 C2 points through one launch of the fused kernel against float64 autograd of the same
    weights: the per-evaluation error that every Newton step of the benched configuration carries."""
    model = pinned_siren(0).to(DEV)
    x = _c2_cloud()[0].to(DEV)
    s, gr = siren.sdf_and_grad(model, x)
    m64 = pinned_siren(0).double().to(DEV)
    es = eg = 0.0
    gmax = 0.0
    for xs, ss, gs in zip(x.split(50_000), s.split(50_000), gr.split(50_000)):
        xx = xs.double().requires_grad_(True)
        s64 = m64(xx).sdf
        g64, = torch.autograd.grad(s64, xx, torch.ones_like(s64))
        es = max(es, float((ss.double() - s64.detach().reshape(-1)).abs().max()))
        eg = max(eg, float((gs.double() - g64).abs().max()))
        gmax = max(gmax, float(g64.abs().max()))
    print("200k rows: sdf err %.2e abs, grad err %.2e abs = %.2e of max |grad| %.1f" % (es, eg, eg / gmax, gmax))
    assert es < 2e-6 and eg < 3e-5 * gmax


def test_projection_of_the_bench_cloud_vs_reference_golden(golden):
    """First 4 096 points of the C2 cloud: fused path vs the reference's own UniformProjection (CPU fp32)."""
    g = golden("pinned_siren")
    model = pinned_siren(0).to(DEV)
    x = torch.as_tensor(g["x"]).to(DEV)
    proj = UniformProjection(proj_max_iters=10, proj_tolerance=5e-5, knn_k=8, sample_iters=1)
    res = proj._project_points(model, x.clone(), torch.tensor([4096], device=DEV))
    mask, want_mask = res.mask[0].cpu(), torch.as_tensor(g["proj_mask"][0])
    agree = mask == want_mask
    both = mask & want_mask
    ok = _within(res.points[0].cpu()[both], torch.as_tensor(g["proj_points"][0])[both])
    nrm_ok = torch.isclose(res.normals[0].cpu()[both][ok], torch.as_tensor(g["proj_normals"][0])[both][ok],
                           rtol=1e-3, atol=1e-3 * float(np.abs(g["proj_normals"]).max()))
    print("4096-point golden: converged %.4f (reference %.4f), mask agreement %.4f, converged rows within 1e-4: "
          "%.4f (excluded %d of %d), their normals within 1e-3: %.4f"
          % (float(mask.float().mean()), float(want_mask.float().mean()), float(agree.float().mean()),
             float(ok.float().mean()), int((~ok).sum()), int(both.sum()), float(nrm_ok.all(-1).float().mean())))
    assert float(agree.float().mean()) > 0.97
    assert float(ok.float().mean()) > 0.97
    assert float(nrm_ok.all(-1).float().mean()) > 0.99
    # end to end (project -> filter -> resample -> re-project) as a point set: rows shift when a mask bit flips
    out = proj.project_points(x.clone(), model, skip_upsampling=True)
    pa = out["levelset_points"][0][out["mask"][0]]
    pb = torch.as_tensor(g["points"][0][g["mask"][0]]).to(DEV)
    assert abs(pa.shape[0] - pb.shape[0]) <= 0.03 * pb.shape[0]
    d = torch.cdist(pa, pb).min(dim=1).values
    print("after resample: %d vs %d valid points, set distance median %.2e, 90%% %.2e"
          % (pa.shape[0], pb.shape[0], float(d.median()), float(d.quantile(0.9))))
    assert float(d.median()) < 1e-4


def test_c2_fused_vs_opaque_per_point_at_200k():
    """The benched configuration itself: 200 000 points x 7 hidden layers, fused vs opaque, per point.
      stage 1  _project_points (10 Newton iterations) on the same cloud;
      stage 2  resample (FRNN K = 9, repulsion, 3-iteration re-projection) from the SAME filtered input
               (the opaque path's survivors), so rows stay aligned;
      stage 3  project_points(skip_upsampling=True) end to end, as counts."""
    fused = pinned_siren(0).to(DEV)
    opaque = pinned_siren(0).as_opaque().to(DEV)
    assert siren.match(fused) is not None and siren.match(opaque) is None
    x = _c2_cloud().to(DEV)
    num = torch.tensor([x.shape[1]], device=DEV)
    with _NoTF32():
        pf = UniformProjection(proj_max_iters=10, proj_tolerance=5e-5, knn_k=8, sample_iters=1)
        po = UniformProjection(proj_max_iters=10, proj_tolerance=5e-5, knn_k=8, sample_iters=1)
        a = pf._project_points(fused, x.clone(), num)
        b = po._project_points(opaque, x.clone(), num)
        agree = (a.mask == b.mask)[0]
        both = (a.mask & b.mask)[0]
        ok = _within(a.points[0][both], b.points[0][both])
        n_both = int(both.sum())
        print("stage 1 (200k x 7 layers, 10 its): converged fused %.4f / opaque %.4f, mask agreement %.5f, "
              "rows converged in both %d, within 1e-4 rel per point: %.5f (excluded %d = %.3f%%)"
              % (float(a.mask.float().mean()), float(b.mask.float().mean()), float(agree.float().mean()), n_both,
                 float(ok.float().mean()), int((~ok).sum()), 100.0 * float((~ok).float().mean())))
        assert float(agree.float().mean()) > 0.97
        assert float(ok.float().mean()) > 0.97
        # the normals returned for those rows (last gradient): direction within 1e-3
        na = torch.nn.functional.normalize(a.normals[0][both][ok], dim=-1)
        nb = torch.nn.functional.normalize(b.normals[0][both][ok], dim=-1)
        cosd = (na * nb).sum(-1)
        print("          normals of those rows: 1 - cos median %.2e, 99%% %.2e"
              % (float((1 - cosd).median()), float((1 - cosd).quantile(0.99))))
        assert float((1 - cosd).quantile(0.99)) < 1e-4

        # stage 2: identical input to both resamplers
        keep = b.mask[0]
        pts = b.points[:, keep].contiguous()
        nrm = b.normals[:, keep].contiguous()
        n2 = torch.tensor([pts.shape[1]], device=DEV)
        ra = pf.resample(fused, pts.clone(), nrm.clone(), n2, sample_iters=1)
        rb = po.resample(opaque, pts.clone(), nrm.clone(), n2, sample_iters=1)
        agree2 = (ra.mask == rb.mask)[0]
        both2 = (ra.mask & rb.mask)[0]
        ok2 = _within(ra.points[0][both2], rb.points[0][both2])
        print("stage 2 (resample of %d rows): valid fused %.4f / opaque %.4f, mask agreement %.5f, within 1e-4 rel "
              "per point: %.5f (excluded %d)" % (pts.shape[1], float(ra.mask.float().mean()),
                                                 float(rb.mask.float().mean()), float(agree2.float().mean()),
                                                 float(ok2.float().mean()), int((~ok2).sum())))
        assert float(agree2.float().mean()) > 0.98
        assert float(ok2.float().mean()) > 0.98

        # stage 3: the public call, both ways
        oa = pf.project_points(x.clone(), fused, skip_upsampling=True)
        ob = po.project_points(x.clone(), opaque, skip_upsampling=True)
    na_, nb_ = int(oa["mask"].sum()), int(ob["mask"].sum())
    print("stage 3 (project_points): %d rows after the filter (opaque %d), valid after resample %d (opaque %d)"
          % (oa["mask"].shape[1], ob["mask"].shape[1], na_, nb_))
    assert abs(oa["mask"].shape[1] - ob["mask"].shape[1]) <= 0.01 * ob["mask"].shape[1]
    assert abs(na_ - nb_) <= 0.01 * nb_


def _project(proj, model, x, run_ahead, **kw):
    from isopoints_b200 import levelset_sampling as ls
    old, ls.RUN_AHEAD = ls.RUN_AHEAD, run_ahead
    try:
        return proj.project_points(x, model, skip_upsampling=True, **kw)
    finally:
        ls.RUN_AHEAD = old


@pytest.mark.parametrize("n,sample_iters", [(200_000, 1), (30_000, 2), (7, 1)])
def test_run_ahead_filter_resample_equals_the_read_back_path_bit_for_bit(n, sample_iters):
    """The survivor count left on the device (filter + resample enqueued at the projection's capacity, one read-back
    at the end) against the count read right after the projection: same kernels on the same live rows, so every
    output -- points, normals, mask, and the cached K-NN table -- has to be identical."""
    model = pinned_siren(0).to(DEV)
    x = _c2_cloud(n).to(DEV)
    pa = UniformProjection(proj_max_iters=10, proj_tolerance=5e-5, knn_k=8, sample_iters=sample_iters)
    pb = UniformProjection(proj_max_iters=10, proj_tolerance=5e-5, knn_k=8, sample_iters=sample_iters)
    a = _project(pa, model, x, True)
    b = _project(pb, model, x, False)
    assert a.keys() == b.keys()
    for k in a:
        assert a[k].shape == b[k].shape and torch.equal(a[k], b[k]), k
    if pa._knn_idx is not None or pb._knn_idx is not None:
        assert torch.equal(pa._knn_idx, pb._knn_idx) and torch.equal(pa._knn_dists, pb._knn_dists)
        assert torch.equal(pa._knn_nn, pb._knn_nn)
    print("n = %d: %d survivors, identical" % (n, a["levelset_points"].shape[1]))


def test_run_ahead_when_nothing_converges_returns_the_unfiltered_projection():
    """:396-399 -- no survivor: the speculative resample runs over an empty cloud (finite grid, no exception) and the
    unfiltered projection comes back without 'levelset_normals', as on the read-back path."""
    model = pinned_siren(0).to(DEV)
    x = _c2_cloud(5_000).to(DEV)
    kw = dict(proj_max_iters=1, proj_tolerance=1e-12, knn_k=8, sample_iters=1)
    a = _project(UniformProjection(**kw), model, x, True)
    b = _project(UniformProjection(**kw), model, x, False)
    assert set(a.keys()) == set(b.keys()) == {"levelset_points", "mask"}
    assert not bool(a["mask"].any())
    assert torch.equal(a["levelset_points"], b["levelset_points"]) and torch.equal(a["mask"], b["mask"])


def test_run_ahead_redoes_the_search_when_the_cell_table_sized_ahead_is_too_small():
    """The run-ahead path sizes the FRNN cell table before it knows the grid; with a budget far below what this cloud
    needs the capped parameters hand out a grid no query reaches (nothing out of bounds), the deferred check sees it,
    and the step is redone with the sizes read back: same result as the read-back path."""
    from isopoints_b200 import frnn
    model = pinned_siren(0).to(DEV)
    x = _c2_cloud(30_000).to(DEV)
    kw = dict(proj_max_iters=10, proj_tolerance=5e-5, knn_k=8, sample_iters=1)
    old, frnn.SPECULATIVE_CELLS = frnn.SPECULATIVE_CELLS, 1000
    try:
        a = _project(UniformProjection(**kw), model, x, True)
    finally:
        frnn.SPECULATIVE_CELLS = old
    b = _project(UniformProjection(**kw), model, x, False)
    for k in b:
        assert torch.equal(a[k], b[k]), k
    assert frnn.DEFERRED_GRID_CHECKS is None


def test_capped_grid_parameters_keep_an_oversized_grid_in_bounds():
    """isob200_frnn_grid_params_capped with a budget below the grid size: g_max reports the true size, the parameters
    describe a single cell, and build + query over a (N, cap) table find nothing (and touch nothing outside it)."""
    from isopoints_b200 import frnn
    g = torch.Generator().manual_seed(3)
    pts = torch.rand(2, 5000, 3, generator=g).to(DEV)
    lens = torch.tensor([5000, 4000], device=DEV)
    r = torch.full((2,), 0.05, device=DEV)
    frnn.DEFERRED_GRID_CHECKS, old_cells = [], frnn.SPECULATIVE_CELLS
    frnn.SPECULATIVE_CELLS = 64
    try:
        dists, idxs, _, grid = frnn.frnn_grid_points(pts, pts, lens, lens, K=8, r=r)
        (gmax, cap), = frnn.DEFERRED_GRID_CHECKS
    finally:
        frnn.DEFERRED_GRID_CHECKS, frnn.SPECULATIVE_CELLS = None, old_cells
    torch.cuda.synchronize()
    assert cap == 64 and int(gmax) > 64 and grid.pc2_grid_off.shape == (2, 64)
    assert bool((grid.grid_params[:, 7] == 1).all()) and bool((idxs == -1).all())
    ref_d, ref_i, _, ref_grid = frnn.frnn_grid_points(pts, pts, lens, lens, K=8, r=r)
    assert int(gmax) == ref_grid.pc2_grid_off.shape[1] and bool((ref_i[0, :, 0] >= 0).all())
