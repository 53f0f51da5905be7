"""ctypes binding of the C ABI declared in ``include/feos_b200.h`` (``libfeos_b200.so``).

This is the only place the Python layer touches native code.  There is deliberately NO fallback:
if the CUDA library is missing or a call fails, an exception is raised.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# FEOS_B200_LIB: alternative build of the same library (A/B measurements of kernel variants)
LIB_PATH = os.environ.get("FEOS_B200_LIB") or os.path.join(_HERE, "libfeos_b200.so")

_vp = ctypes.c_void_p
_i64 = ctypes.c_int64

# name -> argument types; every symbol include/feos_b200.h declares (tests check the export list)
SIGNATURES = {
    "feos_b200_version": ([], ctypes.c_char_p),
    "feos_b200_last_error": ([], ctypes.c_char_p),
    "feos_b200_launch_count": ([], _i64),
    "feos_b200_measure_fp64_peak": ([ctypes.c_int, _vp], ctypes.c_int),
    "feos_dev_compact": ([_vp, _i64, _vp, _vp, _vp, ctypes.c_int, _vp, ctypes.c_int, _vp], ctypes.c_int),
    "feos_dev_pure_order": ([_vp, _vp, _i64, _vp, _vp], ctypes.c_int),
    "feos_dev_pure_vle_fwd": ([_vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp], ctypes.c_int),
    "feos_dev_pure_vle_bwd": ([_vp, _vp, _vp, _vp, _vp, _vp, _vp, ctypes.c_int, _i64, _vp, _vp, _vp], ctypes.c_int),
    "feos_dev_pure_rho_tp_fwd": ([_vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp], ctypes.c_int),
    "feos_dev_pure_rho_tp_bwd": ([_vp, _vp, _vp, _vp, _vp, _vp, _vp, ctypes.c_int, _i64, _vp, _vp, _vp, _vp], ctypes.c_int),
    "feos_dev_pure_state": ([_vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp], ctypes.c_int),
    "feos_dev_pure_state_bwd": ([_vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp], ctypes.c_int),
    "feos_dev_mix_state_bwd": ([_vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp], ctypes.c_int),
    "feos_dev_gc_state_bwd": ([_vp, ctypes.c_int32, _vp, ctypes.c_int32, _vp, ctypes.c_int32, _vp, _vp, _vp, _vp, _i64,
                               _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp], ctypes.c_int),
    "feos_dev_mix_point_order": ([_vp, _vp, _vp, _i64, ctypes.c_int, _vp, _vp], ctypes.c_int),
    "feos_dev_gc_point_order": ([_vp, ctypes.c_int32, _vp, _vp, _i64, ctypes.c_int, _vp, _vp], ctypes.c_int),
    "feos_dev_mix_point_fwd": ([_vp, _vp, _vp, _vp, _vp, _i64, ctypes.c_int, ctypes.c_int, _vp, _vp, _vp, _vp, _vp, _vp], ctypes.c_int),
    "feos_dev_mix_point_bwd": ([_vp, _vp, _vp, _vp, _vp, _vp, _vp, ctypes.c_int, _i64, ctypes.c_int, ctypes.c_int, _vp, _vp, _vp, _vp, _vp], ctypes.c_int),
    "feos_dev_mix_state": ([_vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp], ctypes.c_int),
    "feos_dev_gc_point_fwd": ([_vp, ctypes.c_int32, _vp, ctypes.c_int32, _vp, ctypes.c_int32, _vp, _vp, _vp, _vp, _vp, _i64,
                               ctypes.c_int, ctypes.c_int, _vp, _vp, _vp, _vp, _vp, _vp], ctypes.c_int),
    "feos_dev_gc_point_bwd": ([_vp, ctypes.c_int32, _vp, ctypes.c_int32, _vp, ctypes.c_int32, _vp, _vp, _vp, _vp, _vp,
                               _vp, _vp, ctypes.c_int, _i64, ctypes.c_int, _vp, _vp, _vp, _vp, _vp, _vp], ctypes.c_int),
    "feos_dev_gc_state": ([_vp, ctypes.c_int32, _vp, ctypes.c_int32, _vp, ctypes.c_int32, _vp, _vp, _vp, _vp, _i64,
                           _vp, _vp, _vp, _vp, _vp], ctypes.c_int),
    "feos_host_gc_point": ([_vp, ctypes.c_int32, _vp, ctypes.c_int32, _vp, ctypes.c_int32, _vp, _vp, _vp, _vp, _vp,
                            _i64, ctypes.c_int, _vp, _vp, _vp], ctypes.c_int),
    "feos_host_pcsaft_bubble_point": ([_vp, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp], ctypes.c_int),
    "feos_host_pcsaft_dew_point": ([_vp, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp], ctypes.c_int),
    "feos_host_pcsaft_vapor_pressure": ([_vp, _vp, _i64, _vp, _vp, _vp], ctypes.c_int),
    "feos_host_pcsaft_liquid_density": ([_vp, _vp, _vp, _i64, _vp, _vp, _vp], ctypes.c_int),
}

_lib = None
_FUNCS = {}


class FeosB200Error(RuntimeError):
    pass


def load():
    """Load ``libfeos_b200.so`` (built by ``__graft_entry__.build()``); raises if it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise FeosB200Error(
                f"{LIB_PATH} not found: build the CUDA extension first (python -c 'import __graft_entry__ as g; g.build()'). "
                "feos_torch_b200 has no CPU fallback."
            )
        lib = ctypes.CDLL(LIB_PATH)
        for name, (args, res) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.argtypes = args
            fn.restype = res
        _lib = lib
    return _lib


def call(name, *args):
    """Call an ``int``-returning entry point and raise with the library's error text on failure."""
    fn = _FUNCS.get(name)
    if fn is None:
        fn = _FUNCS[name] = getattr(load(), name)
    rc = fn(*args)
    if rc != 0:
        raise FeosB200Error(f"{name} failed ({rc}): {load().feos_b200_last_error().decode()}")


TIMINGS = None  # set to {} by bench.py: entry point -> [(start_event, end_event), ...]


def timed_call(name, *args):
    """``call`` bracketed by CUDA events on the current torch stream when ``TIMINGS`` is enabled
    (per-kernel shares inside bench.py's timed region); plain ``call`` otherwise."""
    if TIMINGS is None:
        return call(name, *args)
    import torch

    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    call(name, *args)
    e1.record()
    TIMINGS.setdefault(name, []).append((e0, e1))


# Bubble / dew batches of at least this many states are solved in the state order of csrc/order.cuh (states with
# the same code path and neighbouring (T, x) share a warp); FEOS_B200_STATE_ORDER=0 keeps the input order (A/B runs).
ORDER_MIN_STATES = 1 << 14


def state_order_enabled(n):
    return n >= ORDER_MIN_STATES and os.environ.get("FEOS_B200_STATE_ORDER", "1") != "0"


# The order is a property of the data set: a fit loop solves the same (T, x) tensors every iteration, so the
# permutation is kept for the last few data sets, recognised by the IDENTITY of the caller's tensor objects (weak
# references) and their version counters -- storage addresses are not enough: the caching allocator hands the
# address of a freed temperature tensor to the next one, and a permutation built for other temperatures, while
# still giving correct results (any permutation does), no longer aligns the iteration counts.
#
# "Soft" tensors (the per-state PARAMETERS of the pure path) change in place every optimizer step of a fit; the
# key of a state (code path, chain-length bin, T / T_c estimate) moves slowly with them, and a slightly stale
# permutation costs a little lane utilisation, never correctness.  The permutation is therefore kept across their
# version changes and recomputed every ORDER_REFRESH-th change (FEOS_B200_ORDER_REFRESH, 1 = every change), which
# amortises the key kernel + radix sort (about 10 % of a C2 step) to under 1 %.
_ORDER_CACHE = []
_ORDER_CACHE_SIZE = 8
ORDER_REFRESH = max(1, int(os.environ.get("FEOS_B200_ORDER_REFRESH", "16")))


class _OrderEntry:
    __slots__ = ("kind", "refs", "versions", "soft_seen", "changes", "order", "n")


def cached_order(kind, tensors, compute, n=None, soft=()):
    """``compute(out)`` -> int32 order tensor (written into ``out`` if given, else newly allocated; or None), memoised on (kind, identity + version of ``tensors``,
    ``n``); ``soft`` tensors only age the entry; an entry whose order does not have ``n`` rows is never returned."""
    import weakref

    versions = tuple(t._version for t in tensors)
    for k, e in enumerate(_ORDER_CACHE):
        if (e.kind == kind and e.versions == versions and e.n == n and len(e.refs) == len(tensors)
                and all(r() is t for r, t in zip(e.refs, tensors))):
            # soft tensors: a new version OR a new tensor object (parameters recomputed from other leaves every
            # iteration) counts as one change
            if len(soft) != len(e.soft_seen) or any(r() is not t or v != t._version for (r, v), t in zip(e.soft_seen, soft)):
                e.soft_seen = tuple((weakref.ref(t), t._version) for t in soft)
                e.changes += 1
            if e.order is not None and n is not None and e.order.shape[0] != n:
                del _ORDER_CACHE[k]
                break
            if e.changes >= ORDER_REFRESH:
                # recomputed INTO the same tensor (stream-ordered after its last use): a refresh allocates nothing, so
                # the caching allocator's steady state -- and with it the step time -- is not disturbed
                e.order = compute(e.order)
                e.changes = 0
            _ORDER_CACHE.insert(0, _ORDER_CACHE.pop(k))
            return e.order
    _ORDER_CACHE[:] = [e for e in _ORDER_CACHE if all(r() is not None for r in e.refs)]  # tensors that are gone
    e = _OrderEntry()
    e.kind, e.versions, e.n, e.changes = kind, versions, n, 0
    e.refs = tuple(weakref.ref(t) for t in tensors)
    e.soft_seen = tuple((weakref.ref(t), t._version) for t in soft)
    e.order = compute(None)
    _ORDER_CACHE.insert(0, e)
    del _ORDER_CACHE[_ORDER_CACHE_SIZE:]
    return e.order


def measure_fp64_peak(iters=20000):
    out = ctypes.c_double(0.0)
    call("feos_b200_measure_fp64_peak", iters, ctypes.cast(ctypes.byref(out), ctypes.c_void_p))
    return out.value


def launch_count():
    return int(load().feos_b200_launch_count())


def ptr(t):
    """Device (or host) address of a contiguous tensor / numpy array as a plain int (ctypes converts it to the
    ``void*`` of the signature; building a ``c_void_p`` object per argument costs more than the conversion),
    ``None`` -> NULL."""
    if t is None:
        return None
    if hasattr(t, "data_ptr"):
        return t.data_ptr()
    return t.ctypes.data


def ptr_at(t, byte_offset):
    """``ptr(t)`` advanced by ``byte_offset`` bytes (row-chunked launches)."""
    return t.data_ptr() + byte_offset
