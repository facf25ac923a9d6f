"""The oracle (oracle/port.py) pinned against golden vectors produced by the reference's own
code (tests/golden/make_golden.py).  CPU only."""
import numpy as np
import torch

from oracle import port
from tests.helpers import SphereSDF, TinySiren


def test_projection_sphere_matches_reference(golden):
    g = golden("proj_sphere")
    x = torch.as_tensor(g["x"])
    pts, nrm, mask = port.project_points_padded(SphereSDF(), x, [x.shape[1]], proj_max_iters=10,
                                                proj_tolerance=5e-5)
    assert np.array_equal(mask.numpy(), g["mask"])
    np.testing.assert_allclose(pts.numpy(), g["points"], rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(nrm.numpy(), g["normals"], rtol=1e-6, atol=1e-6)


def test_projection_opaque_module_ragged_matches_reference(golden):
    g = golden("proj_siren")
    x = torch.as_tensor(g["x"])
    pts, nrm, mask = port.project_points_padded(TinySiren(seed=3), x, g["num"].tolist(), proj_max_iters=6,
                                                proj_tolerance=5e-5)
    assert np.array_equal(mask.numpy(), g["mask"])
    np.testing.assert_allclose(pts.numpy(), g["points"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(nrm.numpy(), g["normals"], rtol=1e-4, atol=1e-5)
    assert 0.3 < mask.float().mean() < 0.95   # a mix of converged / non-converged rows


def test_resample_matches_reference(golden):
    g = golden("resample_sphere")
    x = torch.as_tensor(g["x"])[0]
    sdf = SphereSDF()
    p, n, v = port.project_points_packed(sdf, x, proj_max_iters=10, proj_tolerance=5e-5)
    assert bool(v.all())
    # the reference ran its brute-force CPU twin (strict < r2) underneath
    bf = lambda pts, K, r: port.frnn_bruteforce(pts[None], pts[None], K=K, r=r, inclusive=False)[0][0]
    rp, rn, rv = port.resample(sdf, p, n, sample_iters=1, knn_k=8, frnn_fn=bf, proj_tolerance=5e-5)
    assert np.array_equal(rv.numpy(), g["mask"][0])
    np.testing.assert_allclose(rp.numpy(), g["points"][0], rtol=1e-5, atol=2e-6)
    np.testing.assert_allclose(rn.numpy(), g["normals"][0], rtol=1e-5, atol=2e-6)
    rp3, rn3, rv3 = port.resample(sdf, p, n, sample_iters=3, knn_k=8, frnn_fn=bf, proj_tolerance=5e-5)
    assert np.array_equal(rv3.numpy(), g["mask3"][0])
    np.testing.assert_allclose(rp3.numpy(), g["points3"][0], rtol=1e-5, atol=5e-6)


def test_frnn_bruteforce_matches_reference(golden):
    g = golden("frnn_bf")
    idx, d = port.frnn_bruteforce(g["p"], g["p"], g["lens"], g["lens"], K=int(g["K"]), r=float(g["r"]),
                                  inclusive=False)
    assert np.array_equal(idx, g["idxs"])
    np.testing.assert_allclose(d, g["dists"], rtol=1e-6, atol=1e-9)
    assert (idx[1, 1200:] == -1).all() and (idx[0, :, 0] == np.arange(1500)).all()
    idx2, d2 = port.frnn_bruteforce(g["p2d"], g["p2d"], K=5, r=0.05, inclusive=False)
    assert np.array_equal(idx2, g["idxs2d"])
    np.testing.assert_allclose(d2, g["dists2d"], rtol=1e-6, atol=1e-9)


def test_frnn_grid_restatement_equals_bruteforce(golden):
    g = golden("frnn_bf")
    p, lens = g["p"][:, :600], np.array([600, 450])
    rs = np.array([0.1, 0.07], np.float32)
    for D in (3, 2):
        pts = np.ascontiguousarray(p[..., :D])
        params, G = port.frnn_grid_params(pts, lens, rs)
        sp, off, sidx = port.frnn_build_grid(pts, lens, params, G)
        idx_g, d_g = port.frnn_grid_query(pts, lens, sp, off, sidx, lens, params, rs, K=6)
        idx_b, d_b = port.frnn_bruteforce(pts, pts, lens, lens, K=6, r=rs, inclusive=True)
        assert np.array_equal(idx_g, idx_b)
        assert np.array_equal(d_g, d_b)
        # structural invariant the reference checks (frnn_validation_2D_simple.py:22-35)
        for n in range(2):
            assert np.array_equal(sp[n, :lens[n]], pts[n][sidx[n, :lens[n]]])


def test_splat_forward_matches_reference_naive_cpu(golden):
    g = golden("splat_naive_cpu")
    S, K = int(g["S"]), int(g["K"])
    idx, zbuf, qv, occ = port.splat_forward(g["points"], g["ellipse"], g["cutoff"], g["radii"], g["first_idx"],
                                            g["num_points"], float(g["depth_merging_thres"]), S, K,
                                            fine_occupancy=False)
    assert np.array_equal(occ, g["occ"])
    same = (idx == g["idx"]).all(-1)
    # fp32 contraction differs between g++ (golden) and nvcc (oracle): allow a handful of
    # cutoff-boundary pixels, everything else must be identical
    assert same.mean() > 0.999, same.mean()
    np.testing.assert_array_equal(zbuf[same], g["zbuf"][same])
    np.testing.assert_allclose(qv[same], g["qvalue"][same], rtol=2e-5, atol=1e-6)
    assert (idx >= 0).any(-1).mean() > 0.5


def test_splat_bin_counts_against_pairs(golden):
    g = golden("splat_naive_cpu")
    S = int(g["S"])
    cnt = port.splat_bin_counts(g["points"], g["radii"], g["first_idx"], g["num_points"], S, 16)
    assert cnt.shape == (2, 3, 3)
    # every point with z >= 0 lands in at least one bin when inside the frame
    z_ok = (g["points"][:, 2] >= 0).sum()
    assert cnt.sum() >= z_ok


def test_wlop_and_upsample_match_reference(golden):
    g = golden("wlop_upsample")
    P = torch.as_tensor(g["P"])[0]
    bf = lambda a, b, K, r: torch.as_tensor(port.frnn_bruteforce(a[None].numpy(), b[None].numpy(), K=K, r=r,
                                                                   inclusive=False)[0][0])
    X = port.wlop(P, torch.as_tensor(g["noise"]), neighborhood_size=16, iters=3, repulsion_mu=0.5, frnn_fn=bf)
    np.testing.assert_allclose(X.numpy(), g["wlop"][0], rtol=1e-4, atol=2e-6)
    up = port.upsample(torch.as_tensor(g["up_in"])[0], 1300, neighborhood_size=16)
    assert up.shape == (1300, 3) and int(g["up_num"][0]) == 1300
    np.testing.assert_allclose(up.numpy(), g["up_pts"][0], rtol=1e-5, atol=1e-6)


def _rel_rows(a, b):
    """max |a - b| relative to the largest magnitude of each row of b (b can cross zero inside a row)."""
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    sc = np.abs(b).max(axis=-1, keepdims=True) if b.ndim > 1 else np.abs(b)
    return float((np.abs(a - b) / (sc + 1e-30)).max())


def test_ewa_point_info_matches_reference(golden):
    """oracle/port.py ewa_* against SurfaceSplatting._get_per_point_info run from the reference tree
    (ragged views, a 5-point cloud on the `< K` branch, a zero normal)."""
    g = golden("ewa_point_info")
    first, num = g["first_idx"].tolist(), g["num_points"].tolist()
    h = port.ewa_vrk_h(torch.as_tensor(g["sq_dists"]), num)
    assert np.array_equal(h.numpy(), g["vrk_h"])
    assert float(h[num[0]]) == np.float32(5e-4)          # 0.5 * 1e-3: the small-cloud branch
    args = (torch.as_tensor(g["points"]), torch.as_tensor(g["normals"]), first, num, torch.as_tensor(g["proj"]),
            torch.as_tensor(g["vrk_h"]), int(g["image_size"]), float(g["antialiasing_sigma"]), float(g["cutoff"]))
    # float64 restatement vs the reference's float32 run.  The reference's own rounding is the tolerance:
    # a d - b^2 of the 2x2 variance cancels for splats seen at a grazing angle (condition number up to ~1e3
    # here), so its float32 results sit up to ~1.5e-4 from the exact value on the worst points and ~1e-7
    # on the typical one.
    r64 = port.ewa_point_params(*args)
    for name, got in zip(("radii", "ellipse", "cutoff_threshold", "scaler"), r64):
        assert _rel_rows(got.numpy(), g[name]) < 5e-4, name
        err = np.abs(got.numpy() - g[name]) / (np.abs(g[name]).max(axis=-1, keepdims=True) if g[name].ndim > 1
                                               else np.abs(g[name]) + 1e-30)
        assert float(np.median(err)) < 1e-6, name
    # the float32 restatement with a different random tangent frame: same results up to rounding
    r32 = port.ewa_point_params(*args, rand=torch.rand(len(g["points"]), 3, generator=torch.Generator().manual_seed(9)),
                                dtype=torch.float32)
    for name, got in zip(("radii", "ellipse", "cutoff_threshold", "scaler"), r32):
        assert _rel_rows(got.numpy(), g[name]) < 5e-4, name
    assert float(r64[3][7]) == 0.0 and float(g["scaler"][7]) == 0.0      # zero normal: S_k = 0, det M_k = 0


def test_renderable_mask_matches_reference(golden):
    g = golden("ewa_point_info")
    first, num = g["filter_first_idx"].tolist(), g["filter_num_points"].tolist()
    pts, nrm = torch.as_tensor(g["filter_points"]), torch.as_tensor(g["filter_normals"])
    w2v, nmat = torch.as_tensor(g["w2v"]), torch.as_tensor(g["nmat"])
    m, kept = port.renderable_mask(pts, nrm, first, num, w2v, None, float(g["znear"]), float(g["zfar"]))
    assert np.array_equal(m.numpy(), g["mask_depth"])
    m, kept = port.renderable_mask(pts, nrm, first, num, w2v, nmat, float(g["znear"]), float(g["zfar"]))
    assert np.array_equal(m.numpy(), g["mask_renderable"])
    assert sum(kept) == int(g["mask_renderable"].sum()) and 0.2 < m.float().mean() < 0.8


def test_sphere_tracing_matches_reference(golden):
    """oracle/port.py sphere_trace against SphereTracing.project_points run from the reference tree."""
    from tests.helpers import SphereSDF as _Sphere
    g = golden("sphere_trace")
    for name, net in (("siren", TinySiren(seed=3)), ("sphere", _Sphere(radius=0.5))):
        pts, sdf, grad, mask = port.sphere_trace(net, torch.as_tensor(g[name + "_ray0"]), torch.as_tensor(g[name + "_dirs"]),
                                                 proj_max_iters=int(g["proj_max_iters"]), proj_tolerance=5e-5)
        assert np.array_equal(mask.numpy(), g[name + "_mask"]), name
        np.testing.assert_allclose(pts.numpy(), g[name + "_points"], rtol=1e-6, atol=1e-6)
        np.testing.assert_allclose(sdf.numpy(), g[name + "_eval"], rtol=1e-5, atol=1e-7)
        assert 0.2 < mask.float().mean() < 0.9, name      # rays that hit and rays that leave the sphere


def test_pinned_siren_is_the_reference_network(golden):
    """tests/helpers.pinned_siren == DSS.models.common.Siren(dim=3, c_dim=0, hidden_size=256, n_layers=7, ...)
    built after torch.manual_seed(0): fingerprint recorded from the reference's own class, and --
    where the reference tree is present -- state_dict equality with the class itself."""
    import os
    from tests.helpers import pinned_siren
    g = golden("pinned_siren")
    m = pinned_siren(0)
    sd = m.state_dict()
    keys = sorted(sd.keys())
    assert keys == [str(k) for k in g["keys"]]
    fp = np.stack([np.array([float(sd[k].double().sum()), float(sd[k].double().abs().sum()),
                             float(sd[k].reshape(-1)[0]), float(sd[k].reshape(-1)[-1])]) for k in keys])
    np.testing.assert_array_equal(fp, g["fingerprint"])
    state = torch.random.get_rng_state()
    pinned_siren(0)
    assert torch.equal(state, torch.random.get_rng_state())        # the global generator is left alone
    from oracle import ref_python
    if os.path.isdir(ref_python.REF):
        import importlib
        ref_python.load()
        common = importlib.import_module("DSS.models.common")
        with torch.random.fork_rng(devices=[]):
            torch.manual_seed(0)
            ref = common.Siren(dim=3, c_dim=0, hidden_size=256, n_layers=7, first_omega_0=30, hidden_omega_0=30,
                               outermost_linear=True)
        rsd = ref.state_dict()
        assert sorted(rsd.keys()) == keys and all(torch.equal(rsd[k], sd[k]) for k in keys)
        from isopoints_b200 import siren
        assert siren.match(ref, require_cuda=False) is not None    # the real class is what the fused path recognises
    # value / gradient in float64 against the reference class's own
    x = torch.as_tensor(g["x"])[0, :512].double().requires_grad_(True)
    s = m.double()(x).sdf
    gr, = torch.autograd.grad(s, x, torch.ones_like(s))
    np.testing.assert_allclose(s.detach().numpy().reshape(-1), g["sdf64"], rtol=0, atol=1e-12)
    np.testing.assert_allclose(gr.numpy(), g["grad64"], rtol=0, atol=1e-10)


def test_oracle_projection_on_the_pinned_siren_matches_reference(golden):
    """C2's SDF at 7 hidden layers: the oracle's Newton loop + resample against the reference's own
    UniformProjection on the first 4096 points of the bench cloud (60 % of the rows converge)."""
    from tests.helpers import pinned_siren
    g = golden("pinned_siren")
    net = pinned_siren(0).as_opaque()
    x = torch.as_tensor(g["x"])[0]
    p, n, v = port.project_points_packed(net, x, proj_max_iters=10, proj_tolerance=5e-5)
    assert 0.5 < float(v.float().mean()) < 0.7
    agree = v.numpy() == g["proj_mask"][0]
    assert agree.mean() > 0.999          # same torch ops: identical up to the rare |sdf| ~ tol row
    both = agree & v.numpy()
    np.testing.assert_allclose(p.numpy()[both], g["proj_points"][0][both], rtol=1e-4, atol=1e-5)
    if agree.all():
        bf = lambda pts, K, r: port.frnn_bruteforce(pts[None], pts[None], K=K, r=r, inclusive=False)[0][0]
        rp, rn, rv = port.resample(net, p[v], n[v], sample_iters=1, knn_k=8, frnn_fn=bf, proj_tolerance=5e-5)
        m = rv.numpy() == g["mask"][0]
        assert m.mean() > 0.995
        np.testing.assert_allclose(rp.numpy()[m & rv.numpy()], g["points"][0][m & rv.numpy()], rtol=1e-4, atol=1e-5)


def test_insert_oracle_matches_reference(golden):
    g = golden("insert")
    bf = lambda a, b, K, r: port.frnn_bruteforce(a[None].numpy(), b[None].numpy(), K=K, r=r, inclusive=False)
    for m, c, n in (("metric", "child", "child_num"), ("metric2", "child2", "child_num2")):
        child, cn = port.insert(torch.as_tensor(g["ref_points"]), torch.as_tensor(g[m]),
                                torch.as_tensor(g["base"][0]), frnn_fn=bf)
        assert cn == int(g[n][0]) and cn > 0
        np.testing.assert_allclose(child.numpy(), g[c][0], rtol=1e-6, atol=1e-7)
    assert int(g["child_num"][0]) != int(g["child_num2"][0])       # the two selection branches (:189 vs :195-197)


def test_resample_uniformly_oracle_matches_reference(golden):
    g = golden("resample_uniformly")
    P = torch.as_tensor(g["P"][0])
    assert np.array_equal(port.fps(P, 600).numpy(), g["fps_idx"])
    bf = lambda a, b, K, r: torch.as_tensor(
        port.frnn_bruteforce(a[None].numpy(), b[None].numpy(), K=K, r=r, inclusive=False)[0][0])
    out = port.resample_uniformly(P, torch.as_tensor(g["noise"]), frnn_fn=bf)
    assert out.shape == (1200, 3) and int(g["num"][0]) == 1200
    np.testing.assert_allclose(out.numpy(), g["points"][0], rtol=1e-5, atol=1e-6)
