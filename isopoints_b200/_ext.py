"""ctypes loader for libisob200.so -- the C ABI declared in include/isob200.h.

There is deliberately NO fallback: if the shared library is missing or a kernel launch fails,
the call raises.  PyTorch is used only for device memory, streams and autograd plumbing.
"""
import ctypes
import os
import threading

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# ISOB200_LIB: another build of the same library (A/B timing of two kernel versions on one box); entry points it
# does not export are skipped instead of failing the load
LIB_PATH = os.environ.get("ISOB200_LIB") or os.path.join(_HERE, "libisob200.so")

_vp = ctypes.c_void_p
_i = ctypes.c_int
_ll = ctypes.c_longlong
_sz = ctypes.c_size_t
_f = ctypes.c_float
_d = ctypes.c_double

# name -> (restype, argtypes); must match include/isob200.h (tests/test_abi.py checks the export list)
_PROTOS = {
    "isob200_last_error": (ctypes.c_char_p, []),
    "isob200_abi_version": (_i, []),
    "isob200_compiled_arch": (_i, []),
    "isob200_launch_count": (_ll, []),
    "isob200_exclusive_scan_ws_bytes": (_sz, [_i, _i]),
    "isob200_exclusive_scan_i32": (_i, [_vp, _vp, _i, _i, _ll, _ll, _vp, _sz, _vp]),
    "isob200_frnn_grid_params": (_i, [_vp, _vp, _vp, _i, _i, _i, _d, _vp, _vp, _vp, _sz, _vp]),
    "isob200_frnn_grid_params_capped": (_i, [_vp, _vp, _vp, _i, _i, _i, _d, _i, _vp, _vp, _vp, _sz, _vp]),
    "isob200_points_bbox": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp, _sz, _vp]),
    "isob200_frnn_insert_points": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "isob200_frnn_counting_sort": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "isob200_frnn_build_ws_bytes": (_sz, [_i, _i, _i]),
    "isob200_frnn_build": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _sz, _vp]),
    "isob200_frnn_find_nbrs": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i,
                                    _vp, _vp, _i, _i, _vp]),
    "isob200_frnn_gather": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp]),
    "isob200_frnn_gather_backward": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp]),
    "isob200_frnn_backward": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp]),
    "isob200_project_step_ws_bytes": (_sz, [_i]),
    "isob200_project_step": (_i, [_vp, _vp, _vp, _vp, _i, _vp, _vp, _vp, _f, _f, _i, _vp, _vp, _vp, _vp, _sz, _vp]),
    "isob200_trace_step": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _vp, _vp, _vp, _f, _f, _f, _f, _i, _vp, _vp, _vp, _vp, _sz,
                                _vp]),
    "isob200_gather_rows3": (_i, [_vp, _vp, _i, _vp, _vp]),
    "isob200_compact_valid": (_i, [_vp, _vp, _vp, _i, _vp, _vp, _vp, _vp, _sz, _vp]),
    "isob200_project_sphere": (_i, [_vp, _vp, _vp, _ll, _f, _f, _f, _i, _vp]),
    "isob200_siren_set_max_ctas": (_i, [_i]),
    "isob200_siren_set_stagger": (_i, [_i, _i]),
    "isob200_siren_blob_bytes": (_sz, [_i]),
    "isob200_siren_pack_ws_bytes": (_sz, []),
    "isob200_siren_scratch_bytes": (_sz, [_i]),
    "isob200_siren_pack": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _f, _f, _i, _i, _vp, _sz, _vp, _sz, _vp]),
    "isob200_siren_sdf_grad": (_i, [_vp, _i, _vp, _vp, _i, _vp, _vp, _vp, _sz, _vp, _i, _vp]),
    "isob200_siren_project_step": (_i, [_vp, _i, _vp, _vp, _i, _vp, _sz, _vp, _vp, _vp, _vp, _f, _f, _i, _vp, _vp,
                                        _vp, _vp]),
    "isob200_siren_sdf": (_i, [_vp, _i, _vp, _vp, _i, _vp, _vp, _sz, _vp]),
    "isob200_siren_trace_step": (_i, [_vp, _i, _vp, _vp, _i, _vp, _sz, _vp, _vp, _vp, _vp, _f, _f, _f, _f, _i, _vp, _vp,
                                      _vp, _vp]),
    "isob200_ray_nearest_point": (_i, [_vp, _i, _vp, _i, _vp, _i, _vp, _vp, _vp, _vp]),
    "isob200_resample_step": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _vp, _i, _i, _i, _i, _vp, _vp]),
    "isob200_normalize_rows3": (_i, [_vp, _ll, _f, _vp, _vp]),
    "isob200_wlop_density": (_i, [_vp, _vp, _i, _i, _vp, _i, _i, _i, _vp, _vp]),
    "isob200_wlop_step": (_i, [_vp, _vp, _vp, _vp, _i, _i, _vp, _vp, _f, _i, _i, _i, _i, _vp, _vp]),
    "isob200_upsample_sparsity": (_i, [_vp, _vp, _f, _vp, _i, _i, _vp, _i, _i, _i, _vp, _vp, _vp]),
    "isob200_fps": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vp]),
    "isob200_fps_ws_floats": (_sz, [_i, _i]),
    "isob200_fps_ws": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _vp, _sz, _vp, _vp]),
    "isob200_splat_ws_bytes": (_sz, [_i, _i]),
    "isob200_splat_record_bytes": (_i, []),
    "isob200_splat_bin": (_i, [_vp, _vp, _vp, _vp, _i, _ll, _ll, _i, _vp, _sz, _vp, _vp]),
    "isob200_splat_forward": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _ll, _ll, _i, _i, _f, _i, _vp, _sz,
                                   _vp, _ll, _vp, _vp, _vp, _vp, _vp]),
    "isob200_splat_forward_fused": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _ll, _ll, _i, _i, _f, _i, _vp, _sz,
                                         _vp, _ll, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _f, _vp, _vp, _vp, _vp]),
    "isob200_splat_bin_counts": (_i, [_vp, _vp, _vp, _vp, _i, _ll, _i, _i, _vp, _vp]),
    "isob200_splat_count_pairs": (_i, [_vp, _vp, _vp, _vp, _ll, _i, _vp, _vp]),
    "isob200_splat_occ_backward_ws_bytes": (_sz, [_i, _i, _i, _ll]),
    "isob200_splat_occ_backward": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _f, _vp, _i, _i, _i, _ll, _i, _vp, _i,
                                        _vp, _sz, _vp]),
    "isob200_splat_search_radius_ws_bytes": (_sz, [_i]),
    "isob200_splat_search_radius": (_i, [_vp, _vp, _vp, _vp, _i, _ll, _f, _vp, _vp, _sz, _vp]),
    "isob200_splat_zbuf_backward": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _i, _vp]),
    "isob200_splat_visibility": (_i, [_vp, _vp, _ll, _i, _ll, _vp, _vp]),
    "isob200_ewa_vrk_h": (_i, [_vp, _vp, _vp, _i, _ll, _i, _ll, _vp, _vp]),
    "isob200_ewa_point_params": (_i, [_vp, _vp, _vp, _i, _ll, _vp, _i, _vp, _f, _f, _vp, _vp, _vp, _vp, _vp]),
    "isob200_renderable_mask": (_i, [_vp, _vp, _vp, _i, _ll, _vp, _vp, _i, _f, _f, _vp, _vp, _vp]),
    "isob200_splat_blend": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _ll, _i, _i, _f, _vp, _vp, _vp]),
    "isob200_splat_blend_backward": (_i, [_vp, _vp, _vp, _ll, _i, _i, _f, _vp, _i, _vp]),
}

ABI_VERSION = 2      # include/isob200.h as of round 2; a stale .so fails the load instead of misbehaving
_LIB = None
_RAW = None
_TLS = threading.local()

# Optional per-entry-point device timing (bench.py): when PROFILE is a dict, every C-ABI call is
# bracketed by CUDA events on the stream it is launched on; PROFILE[name] collects (start, end).
PROFILE = None
_NO_TIMING = ("_ws_bytes", "_ws_floats", "_blob_bytes", "_scratch_bytes", "isob200_siren_set_max_ctas", "isob200_siren_set_stagger", "isob200_last_error", "isob200_abi_version", "isob200_compiled_arch",
              "isob200_launch_count", "isob200_splat_record_bytes")


class _Lib:
    pass


def _wrap(name, fn):
    if name.endswith(_NO_TIMING) or name in _NO_TIMING:
        return fn

    def timed(*args):
        if PROFILE is None:
            return fn(*args)
        st = torch.cuda.ExternalStream(args[-1]) if args[-1] else torch.cuda.current_stream()
        a = torch.cuda.Event(enable_timing=True)
        b = torch.cuda.Event(enable_timing=True)
        a.record(st)
        rc = fn(*args)
        b.record(st)
        PROFILE.setdefault(name, []).append((a, b))
        return rc

    def call(*args):
        # device guard (the reference's torch extensions guard on the tensors' device): kernels launch on the
        # CURRENT device, the stream handed in belongs to the device `stream()` was last asked for on this thread
        dev = getattr(_TLS, "dev", None)
        if dev is not None and dev != torch.cuda.current_device():
            with torch.cuda.device(dev):
                return timed(*args)
        return timed(*args)
    return call


def exported_symbols():
    return sorted(_PROTOS)


def lib():
    """Load (once) and return the ctypes handle; raises ImportError if the library is absent."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "isopoints_b200: %s not found -- build it with `python -m isopoints_b200.build` "
                "(there is no CPU/PyTorch fallback for the CUDA path)" % LIB_PATH)
        global _RAW
        h = ctypes.CDLL(LIB_PATH)
        w = _Lib()
        for name, (res, args) in _PROTOS.items():
            try:
                fn = getattr(h, name)  # AttributeError if the .so is stale
            except AttributeError:
                if os.environ.get("ISOB200_LIB"):
                    continue
                raise
            fn.restype = res
            fn.argtypes = args
            setattr(w, name, _wrap(name, fn))
        if not os.environ.get("ISOB200_LIB") and h.isob200_abi_version() != ABI_VERSION:
            raise ImportError("isopoints_b200: %s has ABI version %d, the Python side expects %d -- rebuild with "
                              "`python -m isopoints_b200.build --force`" % (LIB_PATH, h.isob200_abi_version(), ABI_VERSION))
        _RAW = h
        _LIB = w
    return _LIB


def check(rc):
    if rc != 0:
        msg = lib().isob200_last_error()
        raise RuntimeError("isob200: " + (msg.decode() if msg else "error %d" % rc))


def ptr(t):
    """Raw device pointer of a tensor (None -> NULL)."""
    if t is None:
        return None
    return t.data_ptr()


def stream(device=None):
    """Raw handle of torch's current stream on ``device``; remembers the device for the launch guard."""
    s = torch.cuda.current_stream(device)
    _TLS.dev = s.device.index
    return s.cuda_stream


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise TypeError("for now only cuda version is supported")


def workspace(nbytes, device):
    return torch.empty((max(int(nbytes), 1),), dtype=torch.uint8, device=device)
