"""Fused SDF value + input gradient for SIREN decoders (csrc/siren.cu).

``UniformProjection._compute_sdf_and_grad`` (DSS/models/levelset_sampling.py:142-170) evaluates
the caller's SDF module through autograd: ``model.forward(x).sdf`` then
``autograd.grad(sdf, x, ones)``.  When that module is the reference's ``Siren`` MLP
(DSS/models/common.py:90-165: ``net = Sequential(SineLayer, ..., SineLayer, Linear)``, every
``SineLayer`` being ``sin(omega_0 * linear(x))``, :56-87) with hidden width 256 and no latent
code, the same two quantities come from one tcgen05 kernel instead.  Any other module keeps the
autograd path; nothing here changes what a caller has to pass.

Recognition is structural (duck-typed, like the rest of the package): the class is called
``Siren`` (or sets ``isob200_fused_siren = True``), has a ``net`` Sequential of modules with
``.linear`` + ``.omega_0`` followed by one ``nn.Linear`` whose first output is the sdf.
Set ``ISOB200_FUSED_SIREN=0`` to force the autograd path (parity tests compare the two).
"""
import os
import weakref

import torch
import torch.nn as nn

from . import _ext

HIDDEN = 256
_CACHE = weakref.WeakKeyDictionary()
ENABLED = os.environ.get("ISOB200_FUSED_SIREN", "1") != "0"
# rows / flops handed to the fused kernel since import (bench.py turns them into the roofline line)
STATS = {"calls": 0, "rows": 0, "flops": 0}
# when a list: the sync-free projection loop appends (first-iteration rows, device tensor of the live
# row counts entering iterations 1..) so that bench.py can count the rows actually evaluated
RECORD = None
# when a list: ``sdf_fn.masked`` appends the device counters of the rows it evaluated (bench_trace.py sums them up
# after the timed region)
MASKED_ROWS = None


def resolve_record():
    """Fold RECORD into STATS (reads the device counters: call outside any timed region)."""
    global RECORD
    if RECORD:
        for m, cnt in RECORD:
            # iteration 0 evaluated all m rows (already counted when it ran without a device count)
            STATS["rows"] += sum(int(c) for c in cnt.tolist()[1:-1]) + (int(m.item()) if torch.is_tensor(m) else 0)
        RECORD = []


def algorithmic_flops(n_rows, n_hidden):
    """fp32-equivalent FLOPs of one evaluation: forward + reverse pass, 2 flops per MAC, over
    the first layer (3 x 256), the n_hidden 256 x 256 layers and the 256 x 1 head."""
    return 2 * 2 * n_rows * (3 * HIDDEN + n_hidden * HIDDEN * HIDDEN + HIDDEN)


class _Spec(object):
    __slots__ = ("first", "hidden", "last", "omega0", "omega")


def match(model, forward_kwargs=None, require_cuda=True):
    """Return the layer spec if ``model`` is a fusable SIREN SDF, else None."""
    if not ENABLED or not isinstance(model, nn.Module):
        return None
    if type(model).__name__ != "Siren" and not getattr(model, "isob200_fused_siren", False):
        return None
    if forward_kwargs:
        c = forward_kwargs.get("c", None)
        others = [k for k in forward_kwargs if k != "c"]
        if others or (c is not None and c.numel() > 0):
            return None
    if getattr(model, "use_activation", False):
        return None
    fields = getattr(model, "_out_fields", None)
    if fields is not None and (len(fields) == 0 or fields[0] != "sdf"):
        return None
    net = getattr(model, "net", None)
    if not isinstance(net, nn.Sequential) or len(net) < 3:
        return None
    mods = list(net)
    sines, last = mods[:-1], mods[-1]
    if not isinstance(last, nn.Linear):
        return None
    for m in sines:
        if not isinstance(getattr(m, "linear", None), nn.Linear) or not hasattr(m, "omega_0"):
            return None
        if m._forward_hooks or m._forward_pre_hooks:
            return None
    if model._forward_hooks or model._forward_pre_hooks:
        return None
    first, hidden = sines[0].linear, [m.linear for m in sines[1:]]
    if first.in_features != 3 or first.out_features != HIDDEN or last.in_features != HIDDEN:
        return None
    if any(h.in_features != HIDDEN or h.out_features != HIDDEN for h in hidden):
        return None
    if not (1 <= len(hidden) <= 32):
        return None
    omegas = {float(m.omega_0) for m in sines[1:]}
    if len(omegas) != 1:
        return None
    params = [first.weight, last.weight] + [h.weight for h in hidden]
    if any((require_cuda and not p.is_cuda) or p.dtype != torch.float32 for p in params):
        return None
    s = _Spec()
    s.first, s.hidden, s.last = first, hidden, last
    s.omega0, s.omega = float(sines[0].omega_0), omegas.pop()
    return s


def _version_key(spec):
    ps = [spec.first.weight, spec.first.bias, spec.last.weight, spec.last.bias]
    for h in spec.hidden:
        ps += [h.weight, h.bias]
    return tuple((p.data_ptr(), p._version) if p is not None else None for p in ps) + (spec.omega0, spec.omega)


def _pack(spec):
    lib = _ext.lib()
    if "ISOB200_SIREN_STAGGER" in os.environ:       # tuning knob of csrc/siren.cu: cycles (-1 = half a tile, 0 = off)
        lib.isob200_siren_set_stagger(int(os.environ["ISOB200_SIREN_STAGGER"]), 0)
    dev = spec.first.weight.device
    L = len(spec.hidden)
    with torch.no_grad():
        w0 = spec.first.weight.detach().contiguous()
        b0 = spec.first.bias.detach().contiguous() if spec.first.bias is not None else None
        wh = torch.stack([h.weight.detach() for h in spec.hidden]).contiguous()
        if all(h.bias is not None for h in spec.hidden):
            bh = torch.stack([h.bias.detach() for h in spec.hidden]).contiguous()
        else:
            bh = torch.stack([h.bias.detach() if h.bias is not None else torch.zeros(HIDDEN, device=dev)
                              for h in spec.hidden]).contiguous()
        wl = spec.last.weight.detach()[0].contiguous()
        bl = spec.last.bias.detach()[:1].contiguous() if spec.last.bias is not None else None
    nbytes = lib.isob200_siren_blob_bytes(L)
    blob = torch.empty((nbytes,), dtype=torch.uint8, device=dev)
    ws = _ext.workspace(lib.isob200_siren_pack_ws_bytes(), dev)
    _ext.check(lib.isob200_siren_pack(_ext.ptr(w0), _ext.ptr(b0), _ext.ptr(wh), _ext.ptr(bh), _ext.ptr(wl),
                                      _ext.ptr(bl), spec.omega0, spec.omega, HIDDEN, L, _ext.ptr(blob), nbytes,
                                      _ext.ptr(ws), ws.numel(), _ext.stream(dev)))
    scratch = _ext.workspace(lib.isob200_siren_scratch_bytes(L), dev)
    return blob, scratch, L


def invalidate(model=None):
    """Drop the packed operand images of ``model`` (all models if None).  Only needed after a parameter
    was modified behind autograd's back (``p.data.mul_(...)`` does not bump ``p._version``); optimizer
    steps, ``load_state_dict`` and any in-place op on the parameter itself are detected."""
    if model is None:
        _CACHE.clear()
    else:
        _CACHE.pop(model, None)


def packed(model, spec):
    """(blob, scratch, L) for the current parameter values; re-packed when any parameter changed
    (storage pointer or autograd version counter, see ``invalidate``)."""
    key = _version_key(spec)
    ent = _CACHE.get(model)
    if ent is None or ent[0] != key:
        ent = (key, _pack(spec))
        _CACHE[model] = ent
    return ent[1]


def sdf_and_grad(model, points, forward_kwargs=None, n_dev=None, dbg_gemm=None, spec=None, out=None, pk=None):
    """sdf (n,) and d sdf/d x (n,3) of a fusable SIREN at ``points`` (n,3) fp32 cuda, or None when
    ``model`` is not fusable.  ``n_dev``: optional int32 device scalar with the live row count (rows
    past it are left untouched); ``spec``: result of an earlier ``match``; ``out``: preallocated
    (sdf, grad) buffers with at least n rows; ``pk``: result of an earlier ``packed`` (same parameters)."""
    if spec is None:
        spec = match(model, forward_kwargs)
    if spec is None:
        return None
    _ext.require_cuda(points)
    lib = _ext.lib()
    x = points.reshape(-1, 3)
    if x.dtype != torch.float32:
        return None
    x = x.contiguous()
    n = x.shape[0]
    dev = x.device
    blob, scratch, L = pk if pk is not None else packed(model, spec)
    if out is not None:
        sdf, grad = out[0][:n], out[1][:n]
    else:
        sdf = torch.empty((n,), dtype=torch.float32, device=dev)
        grad = torch.empty((n, 3), dtype=torch.float32, device=dev)
    dbg = None
    if dbg_gemm is not None:
        dbg = torch.zeros((128, HIDDEN), dtype=torch.float32, device=dev)
    if n > 0:
        STATS["calls"] += 1
        if n_dev is None or RECORD is None:
            STATS["rows"] += n
        _ext.check(lib.isob200_siren_sdf_grad(_ext.ptr(x), n, _ext.ptr(n_dev), _ext.ptr(blob), L, _ext.ptr(sdf),
                                              _ext.ptr(grad), _ext.ptr(scratch), scratch.numel(), _ext.ptr(dbg),
                                              -1 if dbg_gemm is None else int(dbg_gemm), _ext.stream(dev)))
    if dbg_gemm is not None:
        return sdf, grad, dbg
    return sdf, grad


def sdf(model, points, forward_kwargs=None, n_dev=None, spec=None, out=None, pk=None):
    """sdf (n,) of a fusable SIREN at ``points`` (n,3) fp32 cuda through the forward half of the fused kernel
    (no gradient: half the tensor work of ``sdf_and_grad``, same bits as its value), or None when ``model`` is
    not fusable.  Arguments as ``sdf_and_grad``."""
    if spec is None:
        spec = match(model, forward_kwargs)
    if spec is None:
        return None
    _ext.require_cuda(points)
    x = points.reshape(-1, 3)
    if x.dtype != torch.float32:
        return None
    x = x.contiguous()
    n = x.shape[0]
    blob, scratch, L = pk if pk is not None else packed(model, spec)
    val = out[:n] if out is not None else torch.empty((n,), dtype=torch.float32, device=x.device)
    if n > 0:
        STATS["calls"] += 1
        STATS["rows"] += n
        _ext.check(_ext.lib().isob200_siren_sdf(_ext.ptr(x), n, _ext.ptr(n_dev), _ext.ptr(blob), L, _ext.ptr(val),
                                                _ext.ptr(scratch), scratch.numel(), _ext.stream(x.device)))
    return val


class sdf_fn:
    """``sdf`` callable for code that takes one (the reference's RayTracing, levelset_sampling.py:831:
    ``sdf=lambda x: model.decode(x).sdf.squeeze(-1)``): evaluates a fusable decoder with the forward-only fused
    kernel, anything else through ``model.forward(x).sdf`` without autograd."""

    any_size = True     # ray_tracing._eval_chunked: no need to chunk the input when the fused kernel evaluates it

    def __init__(self, model, **forward_kwargs):
        self.model, self.forward_kwargs = model, forward_kwargs
        self._tags, self._rows = {}, None

    def fused(self):
        """True when ``model`` is evaluated by the fused kernel (memory per row: 4 bytes out, no activations)."""
        return match(self.model, self.forward_kwargs) is not None

    def masked(self, points, mask):
        """``zeros(R)`` with the sdf at the rows of ``points`` (R,3) where ``mask`` -- the reference's
        ``out[mask] = sdf(points[mask])`` -- WITHOUT a host synchronisation: the masked rows are compacted on the
        device (``isob200_compact_valid``, which also carries their row numbers), the fused kernel evaluates the
        live count it finds in device memory, and the values are scattered back (rows past the count land in a
        spare slot).  None when the decoder is not fusable or the tensors are not on a GPU."""
        if not (points.is_cuda and self.fused()):
            return None
        lib = _ext.lib()
        dev = points.device
        pts = points.reshape(-1, 3)
        if pts.dtype != torch.float32:
            return None
        pts = pts.contiguous()
        R = pts.shape[0]
        out = torch.zeros((R + 1,), dtype=torch.float32, device=dev)
        if R == 0:
            return out[:0]
        spec = match(self.model, self.forward_kwargs)
        blob, scratch, L = packed(self.model, spec)
        key = (R, dev)
        tag = self._tags.get(key)
        if tag is None:     # row numbers as the bit pattern of a float column: they ride through the compaction
            tag = torch.zeros((R, 3), dtype=torch.float32, device=dev)
            tag[:, 0] = torch.arange(R, dtype=torch.int32, device=dev).view(torch.float32)
            self._tags = {key: tag}
            self._rows = torch.arange(R, dtype=torch.int32, device=dev)
        cp = torch.empty((R, 3), dtype=torch.float32, device=dev)
        ct = torch.empty((R, 3), dtype=torch.float32, device=dev)
        cnt = torch.empty((1,), dtype=torch.int32, device=dev)
        ws = _ext.workspace(lib.isob200_project_step_ws_bytes(R), dev)
        st = _ext.stream(dev)
        _ext.check(lib.isob200_compact_valid(_ext.ptr(pts), _ext.ptr(tag), _ext.ptr(mask.reshape(-1).contiguous().view(torch.uint8)),
                                             R, _ext.ptr(cp), _ext.ptr(ct), _ext.ptr(cnt), _ext.ptr(ws), ws.numel(), st))
        val = torch.empty((R,), dtype=torch.float32, device=dev)
        STATS["calls"] += 1
        _ext.check(lib.isob200_siren_sdf(_ext.ptr(cp), R, _ext.ptr(cnt), _ext.ptr(blob), L, _ext.ptr(val),
                                         _ext.ptr(scratch), scratch.numel(), st))
        if MASKED_ROWS is not None:
            MASKED_ROWS.append(cnt)
        idx = torch.where(self._rows < cnt, ct[:, 0].contiguous().view(torch.int32), torch.full_like(self._rows, R))
        out.scatter_(0, idx.long(), val)
        return out[:R]

    def __call__(self, x):
        v = sdf(self.model, x, self.forward_kwargs) if x.is_cuda else None
        if v is None:
            with torch.no_grad():
                v = self.model.forward(x, **self.forward_kwargs).sdf.reshape(-1)
        return v.reshape(x.shape[:-1])
