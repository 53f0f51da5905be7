"""ctypes loader for tests/hostsim/libhostsim.so (CPU build of the kernels' device functions)."""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None


def build(force=False):
    so = os.path.join(HERE, "libhostsim.so")
    src = os.path.join(HERE, "hostsim.cpp")
    csrc = os.path.join(HERE, "..", "..", "feos_torch_b200", "csrc")
    deps = [src] + [os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith(".cuh")]
    if force or not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-DFEOS_HOST_SIM", "-shared", "-fPIC", "-o", so, src])
    return so


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def c64(x):
    return np.ascontiguousarray(np.asarray(x, dtype=np.float64))


def pure_vle(params, T, robust=True):
    params, T = c64(params), c64(T)
    n = T.shape[0]
    rho = np.empty((n, 2)); vp = np.empty(n); rleq = np.empty(n)
    status = np.empty(n, dtype=np.uint8); evals = np.empty(n, dtype=np.int32)
    lib().sim_pure_vle(_p(params), _p(T), ctypes.c_int64(n), ctypes.c_int(int(robust)), _p(rho), _p(vp), _p(rleq), _p(status), _p(evals))
    return rho, vp, rleq, status, evals


def pure_vle_machine(params, T):
    params, T = c64(params), c64(T)
    n = T.shape[0]
    rho = np.empty((n, 2)); vp = np.empty(n); rleq = np.empty(n)
    status = np.empty(n, dtype=np.uint8); evals = np.empty(n, dtype=np.int32)
    lib().sim_pure_vle_machine(_p(params), _p(T), ctypes.c_int64(n), _p(rho), _p(vp), _p(rleq), _p(status), _p(evals))
    return rho, vp, rleq, status, evals


def pure_vle_bwd(params, T, rho, gv=None, gr=None, generic=False):
    params, T, rho = c64(params), c64(T), c64(rho)
    n = T.shape[0]
    grad = np.zeros((n, 9))
    gv = c64(gv) if gv is not None else None
    gr = c64(gr) if gr is not None else None
    (lib().sim_pure_vle_bwd if generic else lib().sim_pure_vle_bwd_fast)(_p(params), _p(T), _p(rho), ctypes.c_int64(n), _p(gv), _p(gr), _p(grad))
    return grad


def pure_rho_tp(params, T, P):
    params, T, P = c64(params), c64(T), c64(P)
    n = T.shape[0]
    rho = np.empty(n); val = np.empty(n)
    status = np.empty(n, dtype=np.uint8); evals = np.empty(n, dtype=np.int32)
    lib().sim_pure_rho_tp(_p(params), _p(T), _p(P), ctypes.c_int64(n), _p(rho), _p(val), _p(status), _p(evals))
    return rho, val, status, evals


def pure_rho_tp_bwd(params, T, P, rho, g, generic=False):
    params, T, P, rho, g = c64(params), c64(T), c64(P), c64(rho), c64(g)
    n = T.shape[0]
    grad = np.zeros((n, 9)); gp = np.zeros(n)
    lib().sim_pure_rho_tp_bwd(_p(params), _p(T), _p(P), _p(rho), ctypes.c_int64(n), _p(g), _p(grad), _p(gp))
    if not generic:
        lib().sim_pure_rho_tp_bwd_fast(_p(params), _p(T), _p(P), _p(rho), ctypes.c_int64(n), _p(g), _p(grad))
    return grad, gp


def pure_state(params, T, rho, generic=False):
    params, T, rho = c64(params), c64(T), c64(rho)
    n = T.shape[0]
    a = np.empty(n); p = np.empty(n); dp = np.empty(n)
    (lib().sim_pure_state_generic if generic else lib().sim_pure_state)(_p(params), _p(T), _p(rho), ctypes.c_int64(n), _p(a), _p(p), _p(dp))
    return a, p, dp


def mix_state(params, kij, T, rho):
    params, kij, T, rho = c64(params), c64(kij), c64(T), c64(rho)
    n = T.shape[0]
    a = np.empty(n); p = np.empty(n); mu = np.empty((n, 2)); v = np.empty((n, 2))
    lib().sim_mix_state(_p(params), _p(kij), _p(T), _p(rho), ctypes.c_int64(n), _p(a), _p(p), _p(mu), _p(v))
    return a, p, mu, v


def mix_point(params, kij, T, z, bubble, robust=True):
    params, kij, T, z = c64(params), c64(kij), c64(T), c64(z)
    n = T.shape[0]
    rho = np.empty((n, 4)); val = np.empty(n)
    status = np.empty(n, dtype=np.uint8); evals = np.empty(n, dtype=np.int32)
    lib().sim_mix_point(_p(params), _p(kij), _p(T), _p(z), ctypes.c_int64(n), ctypes.c_int(int(bubble)),
                        ctypes.c_int(int(robust)), _p(rho), _p(val), _p(status), _p(evals))
    return rho, val, status, evals


def mix_point_bwd(params, kij, T, rho, bubble, g):
    params, kij, T, rho, g = c64(params), c64(kij), c64(T), c64(rho), c64(g)
    n = T.shape[0]
    grad = np.zeros((n, 19))
    lib().sim_mix_point_bwd(_p(params), _p(kij), _p(T), _p(rho), ctypes.c_int64(n), ctypes.c_int(int(bubble)), _p(g), _p(grad))
    return grad


def _gc_args(pairs, seg, kab, idx, phi, T):
    pairs = np.ascontiguousarray(pairs)
    assert lib().sim_gc_pair_size() == pairs.dtype.itemsize, (lib().sim_gc_pair_size(), pairs.dtype.itemsize)
    return pairs, c64(seg), c64(kab), np.ascontiguousarray(idx, dtype=np.int32), (None if phi is None else c64(phi)), c64(T)


def gc_state(pairs, seg, kab, idx, phi, T, rho):
    pairs, seg, kab, idx, phi, T = _gc_args(pairs, seg, kab, idx, phi, T)
    rho = c64(rho)
    n = T.shape[0]
    a = np.empty(n); p = np.empty(n); mu = np.empty((n, 2)); v = np.empty((n, 2))
    lib().sim_gc_state(_p(pairs), _p(seg), _p(kab), _p(idx), _p(phi), _p(T), _p(rho), ctypes.c_int64(n), _p(a), _p(p), _p(mu), _p(v))
    return a, p, mu, v


def gc_point(pairs, seg, kab, idx, phi, T, z, bubble):
    pairs, seg, kab, idx, phi, T = _gc_args(pairs, seg, kab, idx, phi, T)
    z = c64(z)
    n = T.shape[0]
    rho = np.empty((n, 4)); val = np.empty(n)
    status = np.empty(n, dtype=np.uint8); evals = np.empty(n, dtype=np.int32)
    lib().sim_gc_point(_p(pairs), _p(seg), _p(kab), _p(idx), _p(phi), _p(T), _p(z), ctypes.c_int64(n), ctypes.c_int(int(bubble)),
                       _p(rho), _p(val), _p(status), _p(evals))
    return rho, val, status, evals


def gc_point_bwd(pairs, seg, kab, idx, phi, T, rho, bubble, g, generic=False):
    pairs, seg, kab, idx, phi, T = _gc_args(pairs, seg, kab, idx, phi, T)
    rho, g = c64(rho), c64(g)
    n = T.shape[0]
    gs = np.zeros_like(seg); gk = np.zeros_like(kab); gp = np.zeros((n, 2)); gt = np.zeros(n)
    lib().sim_gc_point_bwd(_p(pairs), _p(seg), ctypes.c_int(seg.shape[0]), _p(kab), ctypes.c_int(kab.shape[0]), _p(idx), _p(phi),
                           _p(T), _p(rho), ctypes.c_int64(n), ctypes.c_int(int(bubble)), _p(g), _p(gs), _p(gk), _p(gp), _p(gt),
                           ctypes.c_int(int(generic)))
    return gs, gk, gp, gt


def gc_point_bwd_kab(pairs, seg, kab, idx, phi, T, rho, bubble, g):
    pairs, seg, kab, idx, phi, T = _gc_args(pairs, seg, kab, idx, phi, T)
    rho, g = c64(rho), c64(g)
    n = T.shape[0]
    gk = np.zeros_like(kab)
    lib().sim_gc_point_bwd_kab(_p(pairs), _p(seg), _p(kab), ctypes.c_int(kab.shape[0]), _p(idx), _p(phi), _p(T), _p(rho),
                               ctypes.c_int64(n), ctypes.c_int(int(bubble)), _p(g), _p(gk))
    return gk


def pure_state_bwd(params, T, rho, ga, gp, gdp):
    params, T, rho, ga, gp, gdp = c64(params), c64(T), c64(rho), c64(ga), c64(gp), c64(gdp)
    n = T.shape[0]
    grad = np.zeros((n, 10))
    lib().sim_pure_state_bwd(_p(params), _p(T), _p(rho), ctypes.c_int64(n), _p(ga), _p(gp), _p(gdp), _p(grad))
    return grad


def mix_state_bwd(params, kij, T, rho, gw):
    params, kij, T, rho, gw = c64(params), c64(kij), c64(T), c64(rho), c64(gw)
    n = T.shape[0]
    grad = np.zeros((n, 21))
    lib().sim_mix_state_bwd(_p(params), _p(kij), _p(T), _p(rho), ctypes.c_int64(n), _p(gw), _p(grad))
    return grad


def gc_state_bwd(pairs, seg, kab, idx, phi, T, rho, gw):
    pairs, seg, kab, idx, phi, T = _gc_args(pairs, seg, kab, idx, phi, T)
    rho, gw = c64(rho), c64(gw)
    n = T.shape[0]
    gs = np.zeros_like(seg); gk = np.zeros_like(kab); gp = np.zeros((n, 2)); gt = np.zeros(n); gr = np.zeros((n, 2))
    lib().sim_gc_state_bwd(_p(pairs), _p(seg), ctypes.c_int(seg.shape[0]), _p(kab), ctypes.c_int(kab.shape[0]), _p(idx), _p(phi),
                           _p(T), _p(rho), ctypes.c_int64(n), _p(gw), _p(gs), _p(gk), _p(gp), _p(gt), _p(gr))
    return gs, gk, gp, gt, gr


def mix_base_eval(params, kij, T, rho):
    params, kij, T, rho = c64(params), c64(kij), c64(T), c64(rho)
    n = T.shape[0]
    out = np.zeros((n, 12))
    lib().sim_mix_base_eval(_p(params), _p(kij), _p(T), _p(rho), ctypes.c_int64(n), _p(out))
    return out[:, :6], out[:, 6:]


def gc_base_eval(pairs, seg, kab, idx, phi, T, rho):
    pairs, seg, kab, idx, phi, T = _gc_args(pairs, seg, kab, idx, phi, T)
    rho = c64(rho)
    n = T.shape[0]
    out = np.zeros((n, 12))
    lib().sim_gc_base_eval(_p(pairs), _p(seg), _p(kab), _p(idx), _p(phi), _p(T), _p(rho), ctypes.c_int64(n), _p(out))
    return out[:, :6], out[:, 6:]


def use_point_machine(on):
    """Route sim_mix_point / sim_gc_point through the resumable PointMachine (work-queue kernels) instead of the loop solver."""
    lib().sim_use_point_machine(ctypes.c_int(int(on)))


def mix_eval_pair(params, kij, T, rho):
    params, kij, T, rho = c64(params), c64(kij), c64(T), c64(rho)
    n = T.shape[0]
    out = np.zeros((n, 12))
    lib().sim_mix_eval_pair(_p(params), _p(kij), _p(T), _p(rho), ctypes.c_int64(n), _p(out))
    return out[:, :6], out[:, 6:]


def mix_bwd_closed(on):
    """sim_mix_point_bwd: closed-form adjoints + reverse sweep (what the kernel runs for non-associating states) or the dual passes."""
    lib().sim_mix_bwd_closed(ctypes.c_int(int(on)))


def mix_order_key(params, T, z):
    params, T, z = c64(params), c64(T), c64(z)
    keys = np.empty(T.shape[0], dtype=np.uint64)
    lib().sim_mix_order_key(_p(params), _p(T), _p(z), ctypes.c_int64(T.shape[0]), _p(keys))
    return keys


def gc_order_key(pair, T, z):
    pair = np.ascontiguousarray(pair, dtype=np.int32)
    T, z = c64(T), c64(z)
    keys = np.empty(T.shape[0], dtype=np.uint64)
    lib().sim_gc_order_key(_p(pair), _p(T), _p(z), ctypes.c_int64(T.shape[0]), _p(keys))
    return keys


def order_bits(which):
    return int(lib().sim_order_bits(which))


def pure_order_key(params, T):
    params, T = c64(params), c64(T)
    keys = np.empty(T.shape[0], dtype=np.uint64)
    lib().sim_pure_order_key(_p(params), _p(T), ctypes.c_int64(T.shape[0]), _p(keys))
    return keys, int(lib().sim_pure_order_bits())


def mix_point_bwd_kij(params, kij, T, rho, bubble, g=None):
    """k_ij-only backward (mix_point_bwd_kij_kernel): g * d p / d k_ij per state."""
    params, kij, T, rho = c64(params), c64(kij), c64(T), c64(rho)
    n = T.shape[0]
    g = c64(np.ones(n) if g is None else g)
    grad = np.zeros(n)
    lib().sim_mix_point_bwd_kij(_p(params), _p(kij), _p(T), _p(rho), ctypes.c_int64(n), ctypes.c_int(int(bubble)), _p(g), _p(grad))
    return grad


def mix_point_plain(params, kij, T, z, bubble):
    """Fast path of the plain-pair forward kernel (mix_point_fwd_kernel<true>); status 3 = not a plain pair, 2 = retry list."""
    params, kij, T, z = c64(params), c64(kij), c64(T), c64(z)
    n = T.shape[0]
    rho = np.full((n, 4), np.nan); value = np.full(n, np.nan)
    status = np.empty(n, dtype=np.uint8); evals = np.zeros(n, dtype=np.int32)
    lib().sim_mix_point_plain(_p(params), _p(kij), _p(T), _p(z), ctypes.c_int64(n), ctypes.c_int(int(bubble)), _p(rho), _p(value), _p(status), _p(evals))
    return rho, value, status, evals


def gc_point_slim(pairs, seg, kab, idx, phi, T, z, bubble):
    """Fast path of the slim-class GC forward kernel (gc_point_fwd_kernel<true>); status 3 = not a slim pair."""
    pairs, seg, kab, idx, phi, T = _gc_args(pairs, seg, kab, idx, phi, T)
    z = c64(z)
    n = T.shape[0]
    rho = np.full((n, 4), np.nan); val = np.full(n, np.nan)
    status = np.empty(n, dtype=np.uint8); evals = np.zeros(n, dtype=np.int32)
    lib().sim_gc_point_slim(_p(pairs), _p(seg), _p(kab), _p(idx), _p(phi), _p(T), _p(z), ctypes.c_int64(n), ctypes.c_int(int(bubble)),
                            _p(rho), _p(val), _p(status), _p(evals))
    return rho, val, status, evals


def gc_state_slim(pairs, seg, kab, idx, phi, T, rho):
    """(a, p, mu, v, hessian (h00, h01, h11), slim flag) through the slim model where the pair allows it."""
    pairs, seg, kab, idx, phi, T = _gc_args(pairs, seg, kab, idx, phi, T)
    rho = c64(rho)
    n = T.shape[0]
    a = np.empty(n); p = np.empty(n); mu = np.empty((n, 2)); v = np.empty((n, 2)); h = np.empty((n, 3))
    ok = np.empty(n, dtype=np.uint8)
    lib().sim_gc_state_slim(_p(pairs), _p(seg), _p(kab), _p(idx), _p(phi), _p(T), _p(rho), ctypes.c_int64(n), _p(a), _p(p),
                            _p(mu), _p(v), _p(h), _p(ok))
    return a, p, mu, v, h, ok.astype(bool)
