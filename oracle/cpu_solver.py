"""TEST / BASELINE INFRASTRUCTURE ONLY -- loader of the compiled CPU solver ``oracle/cpu_solver.cpp``.

``build()`` compiles it with ``g++ -O3 -fopenmp`` into ``oracle/_build/liboracle_cpu.so`` (git-ignored, travels
to the GPU box with the snapshot); ``pure_vle`` / ``pure_liquid_density`` have the signatures of the torch
versions in ``oracle/solver.py`` and are checked against them in ``tests/test_oracle.py``.  Only ``tests/``,
``__graft_entry__`` and ``bench.py``'s CPU-baseline / reference arm use this module.
"""
import ctypes
import os
import subprocess

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "cpu_solver.cpp")
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "liboracle_cpu.so")
_lib = None


def build(force=False):
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(SRC):
        os.makedirs(OUT_DIR, exist_ok=True)
        # -ffp-contract=off: the same IEEE operations in the same order on every host
        subprocess.check_call(["g++", "-O3", "-std=c++17", "-fopenmp", "-ffp-contract=off", "-shared", "-fPIC", "-o", LIB, SRC])
    return LIB


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        lib = ctypes.CDLL(LIB)
        vp, i64 = ctypes.c_void_p, ctypes.c_int64
        lib.oracle_cpu_threads.restype = ctypes.c_int
        lib.oracle_pure_vle.argtypes = [vp, vp, i64, vp, vp, vp]
        lib.oracle_pure_liquid_density.argtypes = [vp, vp, vp, i64, vp, vp]
        lib.oracle_pure_helmholtz.argtypes = [vp, vp, vp, i64, vp, vp, vp]
        _lib = lib
    return _lib


def threads():
    return int(load().oracle_cpu_threads())


def _np(t):
    return np.ascontiguousarray(t.detach().cpu().numpy() if isinstance(t, torch.Tensor) else t, dtype=np.float64)


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def pure_vle(params, temperature):
    """(rho_V, rho_L, failed) like oracle.solver.pure_vle; one state per OpenMP task (src/pcsaft.rs:86-92)."""
    P, T = _np(params), _np(temperature)
    n = T.shape[0]
    rv, rl = np.empty(n), np.empty(n)
    failed = np.empty(n, dtype=np.uint8)
    load().oracle_pure_vle(_p(P), _p(T), n, _p(rv), _p(rl), _p(failed))
    return torch.from_numpy(rv), torch.from_numpy(rl), torch.from_numpy(failed.astype(bool))


def pure_liquid_density(params, temperature, pressure_pa):
    """(rho, failed) like oracle.solver.pure_liquid_density (src/pcsaft.rs:105-129)."""
    P, T, pr = _np(params), _np(temperature), _np(pressure_pa)
    n = T.shape[0]
    rho = np.empty(n)
    failed = np.empty(n, dtype=np.uint8)
    load().oracle_pure_liquid_density(_p(P), _p(T), _p(pr), n, _p(rho), _p(failed))
    return torch.from_numpy(rho), torch.from_numpy(failed.astype(bool))


def pure_helmholtz(params, temperature, density):
    P, T, r = _np(params), _np(temperature), _np(density)
    n = T.shape[0]
    a, a1, a2 = np.empty(n), np.empty(n), np.empty(n)
    load().oracle_pure_helmholtz(_p(P), _p(T), _p(r), n, _p(a), _p(a1), _p(a2))
    return a, a1, a2
