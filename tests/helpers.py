"""Shared synthetic inputs for tests, golden generation and bench (seeded, fp32)."""
import types

import numpy as np
import torch
import torch.nn as nn


class SphereSDF(nn.Module):
    """sdf(x) = |x| - radius, returned as an object with a `.sdf` (n,1) field
    (the interface levelset_sampling.py:160 expects)."""

    def __init__(self, radius=1.0, analytic=False):
        super().__init__()
        self.radius = radius
        if analytic:  # lets isopoints_b200 pick its fused built-in kernel
            self.isob200_analytic_sdf = ("sphere", radius)

    def forward(self, x, **kwargs):
        return types.SimpleNamespace(sdf=x.norm(dim=-1, keepdim=True) - self.radius)


class TinySiren(nn.Module):
    """3 sine layers x 32 + linear head, deterministic init; offset so the zero set is a blob."""

    def __init__(self, seed=0, hidden=32, omega=6.0):
        super().__init__()
        g = torch.Generator().manual_seed(seed)
        self.omega = omega
        dims = [3, hidden, hidden, hidden, 1]
        self.lin = nn.ModuleList(nn.Linear(dims[i], dims[i + 1]) for i in range(4))
        with torch.no_grad():
            for i, l in enumerate(self.lin):
                bound = (1.0 / dims[i]) if i == 0 else (np.sqrt(6.0 / dims[i]) / omega)
                l.weight.copy_((torch.rand(l.weight.shape, generator=g) * 2 - 1) * bound)
                l.bias.copy_((torch.rand(l.bias.shape, generator=g) * 2 - 1) * 0.1)

    def forward(self, x, **kwargs):
        h = x
        for l in self.lin[:-1]:
            h = torch.sin(self.omega * l(h))
        out = self.lin[-1](h) + 0.5 * (x.norm(dim=-1, keepdim=True) - 0.6)
        return types.SimpleNamespace(sdf=out)


def make_splat_inputs(n_views, pts_per_view, S, seed, sigma_px=1.5, aniso=True, behind_frac=0.05):
    """Packed screen-space splats at the kernel boundary: xy ~ U(-0.95,0.95),
    z ~ U(1,3) (a few < 0 to exercise the behind-camera reject), ellipse (a,b,c) of a Gaussian
    with std ~ sigma_px pixels (mildly anisotropic/rotated when `aniso`), cutoff = 1, radii =
    the axis-aligned box of the cutoff ellipse."""
    rng = np.random.RandomState(seed)
    if isinstance(pts_per_view, int):
        pts_per_view = [pts_per_view] * n_views
    P = int(sum(pts_per_view))
    xy = rng.uniform(-0.95, 0.95, size=(P, 2))
    z = rng.uniform(1.0, 3.0, size=(P, 1))
    z[rng.uniform(size=(P, 1)) < behind_frac] *= -1.0
    sig = sigma_px * (2.0 / S)
    if aniso:
        s1 = sig * rng.uniform(0.7, 1.3, size=P)
        s2 = sig * rng.uniform(0.7, 1.3, size=P)
        th = rng.uniform(0, np.pi, size=P)
    else:
        s1 = np.full(P, sig); s2 = np.full(P, sig); th = np.zeros(P)
    c_, s_ = np.cos(th), np.sin(th)
    # covariance V = R diag(s1^2, s2^2) R^T ; Q = [dx dy] V^-1 [dx dy]^T = a dx^2 + b dx dy + c dy^2
    v11 = c_ * c_ * s1 ** 2 + s_ * s_ * s2 ** 2
    v22 = s_ * s_ * s1 ** 2 + c_ * c_ * s2 ** 2
    v12 = c_ * s_ * (s1 ** 2 - s2 ** 2)
    det = v11 * v22 - v12 ** 2
    a = v22 / det
    b = -2 * v12 / det
    c = v11 / det
    cutoff = np.ones(P)
    # bbox of {Q <= cutoff}: half extents sqrt(cutoff * V11), sqrt(cutoff * V22)
    radii = np.stack([np.sqrt(cutoff * v11), np.sqrt(cutoff * v22)], 1)
    num = np.asarray(pts_per_view, np.int64)
    first = np.concatenate([[0], np.cumsum(num)[:-1]]).astype(np.int64)
    return dict(points=np.concatenate([xy, z], 1).astype(np.float32),
                ellipse=np.stack([a, b, c], 1).astype(np.float32),
                cutoff=cutoff.astype(np.float32), radii=radii.astype(np.float32),
                first_idx=first, num_points=num)


class SirenSDF(nn.Module):
    """Random-init SIREN SDF of BASELINE config 2 ("8-layer x 256"): architecture and init of
    DSS/models/common.py:56-165 with dim=3, c_dim=0, hidden_size=256, n_layers=7,
    first_omega_0 = hidden_omega_0 = 30, outermost_linear=True -- 8 sine layers + a linear head.
    (User-side model: the SDF is an opaque nn.Module callback on the hot path.)"""

    def __init__(self, hidden=256, n_layers=7, omega=30.0, seed=0):
        super().__init__()
        self.omega = omega
        g = torch.Generator().manual_seed(seed)
        dims = [3] + [hidden] * (n_layers + 1) + [1]
        self.lin = nn.ModuleList(nn.Linear(dims[i], dims[i + 1]) for i in range(len(dims) - 1))
        with torch.no_grad():
            for i, l in enumerate(self.lin):
                fan = dims[i]
                bound = (1.0 / fan) if i == 0 else (np.sqrt(6.0 / fan) / omega)
                l.weight.copy_((torch.rand(l.weight.shape, generator=g) * 2 - 1) * bound)
                l.bias.copy_((torch.rand(l.bias.shape, generator=g) * 2 - 1) / np.sqrt(fan))

    def forward(self, x, **kwargs):
        h = x
        for l in self.lin[:-1]:
            h = torch.sin(self.omega * l(h))
        return types.SimpleNamespace(sdf=self.lin[-1](h))


class SineLayer(nn.Module):
    """sin(omega_0 * linear(x)) with the SIREN init (DSS/models/common.py:56-87)."""

    def __init__(self, dim, out_dim, is_first=False, omega_0=30.0, gen=None):
        super().__init__()
        self.omega_0 = omega_0
        self.linear = nn.Linear(dim, out_dim)
        bound = (1.0 / dim) if is_first else (np.sqrt(6.0 / dim) / omega_0)
        with torch.no_grad():
            self.linear.weight.copy_((torch.rand(self.linear.weight.shape, generator=gen) * 2 - 1) * bound)
            self.linear.bias.copy_((torch.rand(self.linear.bias.shape, generator=gen) * 2 - 1) / np.sqrt(dim))

    def forward(self, x):
        return torch.sin(self.omega_0 * self.linear(x))


class Siren(nn.Module):
    """Structural twin of the reference decoder DSS/models/common.py:90-165 for
    ``Siren(dim=3, c_dim=0, hidden_size, n_layers, out_dims={'sdf': 1}, outermost_linear=True)``:
    ``net = Sequential(SineLayer(first), n_layers x SineLayer, Linear)`` and a forward returning an
    object with ``.sdf`` -- what isopoints_b200.siren recognises and fuses.  Same weights as
    ``SirenSDF(hidden, n_layers, omega, seed)`` layer for layer (same generator order)."""

    def __init__(self, hidden_size=256, n_layers=7, omega=30.0, seed=0):
        super().__init__()
        g = torch.Generator().manual_seed(seed)
        self.dim, self.c_dim = 3, 0
        self._out_fields, self._out_dims = ("sdf",), (1,)
        self.use_activation = False
        net = [SineLayer(3, hidden_size, True, omega, g)]
        net += [SineLayer(hidden_size, hidden_size, False, omega, g) for _ in range(n_layers)]
        head = nn.Linear(hidden_size, 1)
        with torch.no_grad():
            bound = np.sqrt(6.0 / hidden_size) / omega
            head.weight.copy_((torch.rand(head.weight.shape, generator=g) * 2 - 1) * bound)
            head.bias.copy_((torch.rand(head.bias.shape, generator=g) * 2 - 1) / np.sqrt(hidden_size))
        net.append(head)
        self.net = nn.Sequential(*net)

    def forward(self, coords, c=None, **kwargs):
        return types.SimpleNamespace(sdf=self.net(coords))

    def as_opaque(self):
        """The same weights behind an opaque module (``SirenSDF`` structure: not recognised by the fused path)."""
        n_layers = len(self.net) - 2
        o = SirenSDF(self.net[0].linear.out_features, n_layers, float(self.net[0].omega_0), seed=0)
        with torch.no_grad():
            for dst, src in zip(o.lin, [m.linear for m in list(self.net)[:-1]] + [self.net[-1]]):
                dst.weight.copy_(src.weight)
                dst.bias.copy_(src.bias)
        return o


def pinned_siren(seed=0, hidden_size=256, n_layers=7, omega=30.0):
    """The C2 SDF exactly as SURVEY 8d pins it: ``DSS.models.common.Siren(dim=3, c_dim=0, hidden_size=256,
    n_layers=7, first_omega_0=30, hidden_omega_0=30, outermost_linear=True)`` constructed right after
    ``torch.manual_seed(seed)`` -- i.e. every layer is an ``nn.Linear`` with PyTorch's default init (which draws
    from the global generator: kaiming-uniform weight, then uniform bias) whose weight is then re-drawn with
    ``uniform_`` (first layer +-1/dim, others +-sqrt(6/dim)/omega; DSS/models/common.py:72-127), in construction
    order.  Returned in the ``Siren`` structural twin above; ``tests/test_oracle_golden.py`` asserts state_dict
    equality with the reference's own class.  The global RNG state is left untouched (forked)."""
    with torch.random.fork_rng(devices=[]):
        m = Siren(hidden_size, n_layers, omega, seed=0)   # (nn.Linear's default init draws from the global generator)
        torch.manual_seed(seed)
        dims = [(3, hidden_size)] + [(hidden_size, hidden_size)] * n_layers
        with torch.no_grad():
            for i, (din, dout) in enumerate(dims):
                lin = nn.Linear(din, dout)
                if i == 0:
                    lin.weight.uniform_(-1 / din, 1 / din)
                else:
                    lin.weight.uniform_(-np.sqrt(6 / din) / omega, np.sqrt(6 / din) / omega)
                m.net[i].linear.weight.copy_(lin.weight)
                m.net[i].linear.bias.copy_(lin.bias)
            head = nn.Linear(hidden_size, 1)
            head.weight.uniform_(-np.sqrt(6 / hidden_size) / omega, np.sqrt(6 / hidden_size) / omega)
            m.net[-1].weight.copy_(head.weight)
            m.net[-1].bias.copy_(head.bias)
    return m


def make_cameras(n_views, seed, dist=(2.2, 3.2), focal=2.0, znear=1.0, zfar=100.0):
    """Look-at perspective cameras around the origin in pytorch3d's row-vector convention
    (p_hom @ M).  Returns float32 tensors: w2v (B,4,4) world-to-view, proj (B,4,4) full projection
    (world -> NDC, w = view-space z), nmat (B,3,3) = inverse(w2v)[:3,:3]^T, the matrix normals are
    right-multiplied by (Transform3d.transform_normals)."""
    rng = np.random.RandomState(seed)
    w2v = np.zeros((n_views, 4, 4))
    proj = np.zeros((n_views, 4, 4))
    for b in range(n_views):
        eye = rng.normal(size=3)
        eye *= rng.uniform(*dist) / np.linalg.norm(eye)
        zax = -eye / np.linalg.norm(eye)                     # camera looks at the origin along +z
        up = np.array([0.0, 1.0, 0.0]) if abs(zax[1]) < 0.9 else np.array([1.0, 0.0, 0.0])
        xax = np.cross(up, zax); xax /= np.linalg.norm(xax)
        yax = np.cross(zax, xax)
        R = np.stack([xax, yax, zax], 1)                     # p_view = p_world @ R + T
        T = -eye @ R
        w2v[b, :3, :3] = R
        w2v[b, 3, :3] = T
        w2v[b, 3, 3] = 1.0
        K = np.zeros((4, 4))
        K[0, 0] = K[1, 1] = focal * rng.uniform(0.9, 1.1)
        K[2, 2] = zfar / (zfar - znear)
        K[3, 2] = -zfar * znear / (zfar - znear)
        K[2, 3] = 1.0
        proj[b] = w2v[b] @ K
    nmat = np.linalg.inv(w2v)[:, :3, :3].transpose(0, 2, 1)
    f = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))  # noqa: E731
    return f(w2v), f(proj), f(nmat)


def make_surface_points(pts_per_view, seed, noise=0.02):
    """Packed points near the unit-radius-0.8 sphere with outward normals (a few flipped / zero)."""
    rng = np.random.RandomState(seed)
    P = int(sum(pts_per_view))
    d = rng.normal(size=(P, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    pts = 0.8 * d + noise * rng.normal(size=(P, 3))
    nrm = d + 0.1 * rng.normal(size=(P, 3))
    nrm *= rng.uniform(0.5, 2.0, size=(P, 1))                # not unit length on purpose
    num = np.asarray(pts_per_view, np.int64)
    first = np.concatenate([[0], np.cumsum(num)[:-1]]).astype(np.int64)
    return (torch.from_numpy(pts.astype(np.float32)), torch.from_numpy(nrm.astype(np.float32)),
            torch.from_numpy(first), torch.from_numpy(num))


def make_rays(n, seed, start_radius=1.0, target_radius=0.9):
    """Rays from a sphere of radius ``start_radius`` towards random points within ``target_radius`` of the
    origin (unit directions): most hit a blob around the origin, some graze it or leave the sphere."""
    g = torch.Generator().manual_seed(seed)
    o = torch.randn(n, 3, generator=g)
    o = o / o.norm(dim=-1, keepdim=True) * start_radius
    t = torch.randn(n, 3, generator=g)
    t = t / t.norm(dim=-1, keepdim=True) * target_radius * torch.rand(n, 1, generator=g) ** (1.0 / 3)
    d = t - o
    return o.contiguous(), (d / d.norm(dim=-1, keepdim=True)).contiguous()


def make_camera_rays(n_cams, n_pixels, seed, cam_dist=2.5, target_radius=1.15):
    """``n_cams`` pinhole origins at distance ``cam_dist`` (outside the unit bounding sphere), ``n_pixels`` unit
    rays each towards random points within ``target_radius`` of the origin: most cross the unit sphere, some
    miss it.  Returns cam_loc (B,3), ray_directions (B,P,3)."""
    g = torch.Generator().manual_seed(seed)
    o = torch.randn(n_cams, 3, generator=g)
    o = o / o.norm(dim=-1, keepdim=True) * cam_dist
    t = torch.randn(n_cams, n_pixels, 3, generator=g)
    t = t / t.norm(dim=-1, keepdim=True) * target_radius * torch.rand(n_cams, n_pixels, 1, generator=g) ** (1.0 / 3)
    d = t - o[:, None, :]
    return o.contiguous(), (d / d.norm(dim=-1, keepdim=True)).contiguous()


class PinholeCameras:
    """Minimal perspective cameras with the slice of pytorch3d's camera API the reference's sampling code calls
    (row-vector convention: p_view = p_world @ R + T; NDC xy = focal * xy_view / z_view).  R (B,3,3), T (B,3)."""

    def __init__(self, R, T, focal=2.0, znear=1.0, zfar=100.0):
        self.R, self.T, self.focal, self.znear, self.zfar = R, T, focal, znear, zfar

    @classmethod
    def look_at_origin(cls, n_views, seed, dist=(2.2, 3.2), focal=2.0, device="cpu"):
        w2v, _, _ = make_cameras(n_views, seed, dist=dist, focal=focal)
        return cls(w2v[:, :3, :3].contiguous().to(device), w2v[:, 3, :3].contiguous().to(device), focal)

    def clone(self):
        return PinholeCameras(self.R.clone(), self.T.clone(), self.focal, self.znear, self.zfar)

    def get_camera_center(self):
        return -torch.bmm(self.T[:, None, :], self.R.transpose(1, 2))[:, 0]

    def _w2v(self):
        m = torch.zeros(self.R.shape[0], 4, 4, dtype=self.R.dtype, device=self.R.device)
        m[:, :3, :3], m[:, 3, :3], m[:, 3, 3] = self.R, self.T, 1.0
        return m

    def _proj(self):
        k = torch.zeros(4, 4, dtype=self.R.dtype, device=self.R.device)
        k[0, 0] = k[1, 1] = self.focal
        k[2, 2] = self.zfar / (self.zfar - self.znear)
        k[3, 2] = -self.zfar * self.znear / (self.zfar - self.znear)
        k[2, 3] = 1.0
        return self._w2v() @ k

    def get_world_to_view_transform(self):
        return types.SimpleNamespace(get_matrix=self._w2v)

    def get_full_projection_transform(self):
        return types.SimpleNamespace(get_matrix=self._proj)

    def transform_points(self, points, eps=None):
        hom = torch.cat([points, torch.ones_like(points[..., :1])], dim=-1)
        out = hom @ self._proj()
        return out[..., :3] / out[..., 3:]

    def unproject_points(self, xy_depth, scaled_depth_input=False, world_coordinates=True):
        z = xy_depth[..., 2:3]
        view = torch.cat([xy_depth[..., :2] * z / self.focal, z], dim=-1)
        return (view - self.T[:, None, :]) @ self.R.transpose(1, 2)


def offsurface_inputs(seed=41, n_views=2, n_pix=1200, n_iso=3000, S=64):
    """Inputs shared by the generator and the parity test: cameras, NDC pixels, a disc mask image, and the
    front / back halves (as seen from each camera) of a noisy sphere of iso-points of radius 0.6."""
    g = torch.Generator().manual_seed(seed)
    cams = PinholeCameras.look_at_origin(n_views, seed=seed, focal=2.0)
    pixels = torch.rand(n_views, n_pix, 2, generator=g) * 1.6 - 0.8
    yy, xx = torch.meshgrid(torch.linspace(-1, 1, S), torch.linspace(-1, 1, S), indexing="ij")
    mask_img = ((xx ** 2 + yy ** 2) < 0.33 ** 2).float().expand(n_views, 1, S, S).contiguous()
    d = torch.randn(n_iso, 3, generator=g)
    iso = 0.6 * d / d.norm(dim=-1, keepdim=True) + 0.01 * torch.randn(n_iso, 3, generator=g)
    centre = cams.get_camera_center()
    facing = [((iso * centre[b]).sum(-1) > 0.05) for b in range(n_views)]
    away = [((iso * centre[b]).sum(-1) < -0.05) for b in range(n_views)]
    frontal = [iso[m] for m in facing]
    occluded = [iso[m] for m in away]
    iso_pcl = [iso[: n_iso // 2] * 1.5, iso[n_iso // 2: n_iso // 2 + 700] * 1.5]    # some project outside the disc
    return cams, pixels, mask_img, frontal, occluded, iso_pcl
