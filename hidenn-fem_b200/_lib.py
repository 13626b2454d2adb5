"""ctypes binding of include/hidenn_b200.h (the C-ABI of the CUDA library).

There is deliberately no fallback: if the shared library is missing, or no CUDA device is
visible when a compute entry point is needed, loading raises.
"""
from __future__ import annotations

import ctypes as C
import functools
import os
import re

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("HIDENN_LIB") or os.path.join(HERE, "libhidenn_b200.so")      # HIDENN_LIB: A/B builds (profiles/)
HEADERS = [os.path.join(os.path.dirname(HERE), "include", h) for h in ("hidenn_b200.h", "hidenn_b200_grid.h")]

_lib = None

c_void_p, c_int, c_i64 = C.c_void_p, C.c_int, C.c_int64


class HidennError(RuntimeError):
    pass


def declared_symbols():
    """Every function the public headers declare (used by the CPU export test)."""
    names = []
    for h in HEADERS:
        if not os.path.exists(h):
            continue
        txt = open(h).read()
        txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
        names += re.findall(r"\b(hidenn_[a-z0-9_]+)\s*\(", txt)
    return sorted(set(names))


def lib():
    """Load (building in-tree if sources are newer) and return the ctypes handle."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH) or os.environ.get("HIDENN_REBUILD"):
        from . import build as _build
        _build.build()
    if not os.path.exists(LIB_PATH):
        raise HidennError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'`. "
                          "There is no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    L.hidenn_last_error.restype = C.c_char_p
    for name in declared_symbols():
        if name in ("hidenn_last_error", "hidenn_tri_plan_destroy"):
            continue
        getattr(L, name).restype = c_int
    for name in ("hidenn_1d_scratch_size", "hidenn_q1_l2_partials", "hidenn_halo_p2p_bytes", "hidenn_1d_bar_step_scratch"):
        getattr(L, name).restype = C.c_int64
    L.hidenn_tri_plan_destroy.restype = None
    L.hidenn_tri_plan_destroy.argtypes = [c_void_p]
    _lib = L
    return L


def check(rc, what=""):
    if rc != 0:
        msg = lib().hidenn_last_error().decode()
        raise HidennError(f"{what}: {msg}" if what else msg)


def require_cuda():
    n = lib().hidenn_device_count()
    if n <= 0:
        raise HidennError("no CUDA device visible: the HiDeNN B200 path has no CPU fallback "
                          f"({lib().hidenn_last_error().decode()})")
    return n


def ptr(t):
    """Raw pointer of a tensor (None -> NULL)."""
    if t is None:
        return c_void_p(0)
    return c_void_p(t.data_ptr())


def stream_ptr(device=None):
    """cudaStream_t of torch's current stream on the current device (raw handle, no Stream object)."""
    if device is None:
        idx = torch.cuda.current_device()
    else:
        idx = torch.device(device).index
        if idx is None:
            idx = torch.cuda.current_device()
    return c_void_p(torch._C._cuda_getCurrentRawStream(idx))


def on_device(f):
    """Decorator for autograd.Function.forward / backward: run the body with the CUDA current device set to the
    device of the first CUDA tensor argument, so `stream_ptr()`, allocations and the C-ABI launches all bind to the
    tensors' device even when the caller's current device is another one (the reference's plain torch ops do)."""
    @functools.wraps(f)
    def wrapper(ctx, *args):
        for t in args:
            if isinstance(t, torch.Tensor) and t.is_cuda:
                if t.device.index != torch.cuda.current_device():
                    with torch.cuda.device(t.device):
                        return f(ctx, *args)
                break
        return f(ctx, *args)
    return wrapper


def check_ids(ids, n, what):
    """Index semantics of torch advanced indexing for user-supplied element / edge ids: negative ids wrap, ids
    outside [-n, n) raise IndexError (the kernels index raw tables).  Skipped while a CUDA graph is being captured."""
    if ids.numel() == 0 or torch.cuda.is_current_stream_capturing():
        return ids
    lo, hi = torch.aminmax(ids)
    lo, hi = int(lo), int(hi)
    if lo < -n or hi >= n:
        raise IndexError(f"{what}: index {lo if lo < -n else hi} is out of bounds for dimension 0 with size {n}")
    if lo < 0:
        ids = torch.where(ids < 0, ids + n, ids)
    return ids


def suffix(dtype):
    if dtype == torch.float64:
        return "f64"
    if dtype == torch.float32:
        return "f32"
    raise HidennError(f"unsupported dtype {dtype}: the kernels are FP64 and FP32")


class _NullRange:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


class _NvtxRange:
    def __init__(self, name):
        self.name = name

    def __enter__(self):
        torch.cuda.nvtx.range_push(self.name)
        return self

    def __exit__(self, *a):
        torch.cuda.nvtx.range_pop()
        return False


_NVTX = bool(os.environ.get("HIDENN_NVTX"))
_null_range = _NullRange()


def nvtx(name):
    """NVTX range around a fused call when HIDENN_NVTX=1 (tracing hook of SURVEY §5); free otherwise."""
    return _NvtxRange(name) if _NVTX else _null_range


_fn_cache = {}


def fn(name, dtype):
    key = (name, dtype)
    f = _fn_cache.get(key)
    if f is None:
        f = getattr(lib(), f"{name}_{suffix(dtype)}")
        _fn_cache[key] = f
    return f
