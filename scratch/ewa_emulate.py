"""numpy float32 transcription of csrc/ewa.cu point_params_kernel, checked against the fp64 oracle (CPU)."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from oracle import port

f = np.float32
if len(sys.argv) > 1:
    from tests.helpers import make_cameras, make_surface_points
    views = [40000] * 8
    pts, nrm, first, num = (x.numpy() for x in make_surface_points(views, seed=len(views)))
    proj = make_cameras(8, seed=7)[1].numpy()
    h = (torch.rand(len(pts), generator=torch.Generator().manual_seed(1)) * 2e-3 + 5e-5).numpy()
    g = dict(antialiasing_sigma=float(sys.argv[2]), image_size=int(sys.argv[1]), cutoff=float(sys.argv[3]))
else:
    g = np.load("tests/golden/ewa_point_info.npz")
    pts, nrm, proj, h = g["points"], g["normals"], g["proj"], g["vrk_h"]
    first, num = g["first_idx"], g["num_points"]
b = np.repeat(np.arange(len(num)), num)
M = proj[b]
x, y, z = pts[:, 0], pts[:, 1], pts[:, 2]
pv = f(float(g["antialiasing_sigma"]) * (2.0 / int(g["image_size"])) ** 2)
cut = f(g["cutoff"])
EPS = f(1e-17)
ed = lambda v: np.where(v < 0, f(-1), f(1)) * np.maximum(np.abs(v), EPS)
xv = x * M[:, 0, 0] + y * M[:, 1, 0] + z * M[:, 2, 0] + M[:, 3, 0]
yv = x * M[:, 0, 1] + y * M[:, 1, 1] + z * M[:, 2, 1] + M[:, 3, 1]
t = x * M[:, 0, 3] + y * M[:, 1, 3] + z * M[:, 2, 3] + M[:, 3, 3]
t2 = ed(t * t)
it = f(1) / ed(t)
j30, j31 = (f(-1) / t2) * xv, (f(-1) / t2) * yv
w0 = M[:, :3, 0] * it[:, None] + M[:, :3, 3] * j30[:, None]
w1 = M[:, :3, 1] * it[:, None] + M[:, :3, 3] * j31[:, None]
a = np.abs(nrm)
sel = np.where((a[:, 0] <= a[:, 1]) & (a[:, 0] <= a[:, 2]), 0, np.where(a[:, 1] <= a[:, 2], 1, 2))
e = np.eye(3, dtype=f)[sel]
u0 = np.cross(nrm, e).astype(f)
u0 = u0 / np.maximum(np.linalg.norm(u0, axis=1, keepdims=True), f(1e-12))
u1 = np.cross(nrm, u0).astype(f)
u1 = u1 / np.maximum(np.linalg.norm(u1, axis=1, keepdims=True), f(1e-12))
m00, m01 = (u0 * w0).sum(1), (u0 * w1).sum(1)
m10, m11 = (u1 * w0).sum(1), (u1 * w1).sum(1)
dmk = m00 * m11 - m01 * m10
v00, v01, v11 = h * (m00 * m00 + m10 * m10), h * (m00 * m01 + m10 * m11), h * (m01 * m01 + m11 * m11)
A, D = v00 + pv, v11 + pv
hd = h * dmk
det = hd * hd + pv * (v00 + v11) + pv * pv
idet = f(1) / det
ea, eb, ec = D * idet, f(-2) * v01 * idet, A * idet
den = ed(f(4) * idet)
ry = np.sqrt(np.maximum(np.abs(f(4) * ea * cut / den), EPS))
rx = np.sqrt(np.maximum(np.abs(f(4) * ec * cut / den), EPS))
sk = np.abs(dmk) / ed(np.sqrt(np.maximum(np.abs(det * f(39.478417604357434)), EPS)))
for arr in (rx, ea, sk):
    assert arr.dtype == np.float32
r64 = port.ewa_point_params(torch.as_tensor(pts), torch.as_tensor(nrm), first.tolist(), num.tolist(),
                            torch.as_tensor(proj), torch.as_tensor(h), int(g["image_size"]),
                            float(g["antialiasing_sigma"]), float(g["cutoff"]))
def rel(a, b):
    b = b.numpy()
    sc = np.abs(b).max(-1, keepdims=True) if b.ndim > 1 else np.abs(b)
    return float((np.abs(a - b) / (sc + 1e-30)).max())
print("radii", rel(np.stack([rx, ry], 1), r64[0]), "ellipse", rel(np.stack([ea, eb, ec], 1), r64[1]), "scaler", rel(sk, r64[3]))
if len(sys.argv) > 1:
    sfl = np.abs(sk - r64[3].numpy()) / (np.abs(r64[3].numpy()) + 1e-6 * np.abs(r64[3].numpy()).max())
    print("scaler with floor", sfl.max())
    ae = np.abs(sk - r64[3].numpy()); s64 = np.abs(r64[3].numpy())
    print("scaler abs err max", ae.max(), "median s", np.median(s64), "max s", s64.max(), "max (ae - 1e-4 s)/median", ((ae - 1e-4 * s64) / np.median(s64)).max())
    front = np.linalg.norm(np.cross(w0, w1), axis=1) / (2 * np.pi * np.sqrt(det))
    print("err relative to the frontal scaler", (ae / front).max())
    e = np.stack([ea, eb, ec], 1).astype(np.float64); r = np.stack([rx, ry], 1).astype(np.float64)
    ident = r[:, 0] ** 2 * (4 * e[:, 0] * e[:, 2] - e[:, 1] ** 2) / (4 * e[:, 2])
    print("identity", np.abs(ident / float(g["cutoff"]) - 1).max())
    sys.exit()
print("vs golden: radii", rel(np.stack([rx, ry], 1), torch.as_tensor(g["radii"])), "ellipse",
      rel(np.stack([ea, eb, ec], 1), torch.as_tensor(g["ellipse"])), "scaler", rel(sk, torch.as_tensor(g["scaler"])))
s64 = r64[3].numpy()
err = np.abs(sk - s64)
i = np.argmax(err / (np.abs(s64) + 1e-30))
print("worst scaler point", i, "value", s64[i], "median", np.median(s64), "abs err", err[i], "err/max", err.max() / s64.max())
n_hat = nrm / np.maximum(np.linalg.norm(nrm, axis=1, keepdims=True), 1e-12)
tp = (n_hat * np.cross(w0, w1)).sum(1).astype(f)
print("triple product variant rel err", float((np.abs(np.abs(tp) / ed(np.sqrt(np.maximum(np.abs(det * f(39.478417604357434)), EPS))) - s64) / (np.abs(s64) + 1e-30)).max()))
