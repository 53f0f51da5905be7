"""numpy-in / numpy-out classes with the exact signatures of the reference's native module
``feos_torch.feos_torch`` (Rust/PyO3: `/root/reference/src/pcsaft.rs:13-80`,
`src/gc_pcsaft.rs:15-100`), backed by the host-buffer entry points of ``include/feos_b200.h``."""
import ctypes

import numpy as np

from . import _lib


def _f64(x):
    return np.ascontiguousarray(x, dtype=np.float64)


class PcSaft:
    @staticmethod
    def vapor_pressure(parameters, temperature):
        """-> (rho f64[n_ok,4] = (rho_V, rho_L, 0, 0) in 1/A^3, failed bool[n]); src/pcsaft.rs:19-26."""
        p, t = _f64(parameters), _f64(temperature)
        n = t.shape[0]
        if p.shape != (n, 8):
            raise ValueError("parameters must have shape [n, 8]")
        rho = np.empty((n, 4), dtype=np.float64)
        status = np.empty((n,), dtype=np.uint8)
        n_ok = ctypes.c_int64(0)
        _lib.call("feos_host_pcsaft_vapor_pressure", _lib.ptr(p), _lib.ptr(t), n, _lib.ptr(rho), _lib.ptr(status),
                  ctypes.cast(ctypes.byref(n_ok), ctypes.c_void_p))
        return rho[: n_ok.value].copy(), status.astype(bool)

    @staticmethod
    def liquid_density(parameters, temperature, pressure):
        """-> (rho f64[n_ok] in 1/A^3, failed bool[n]); src/pcsaft.rs:29-41."""
        p, t, pr = _f64(parameters), _f64(temperature), _f64(pressure)
        n = t.shape[0]
        if p.shape != (n, 8) or pr.shape != (n,):
            raise ValueError("parameters must have shape [n, 8], pressure [n]")
        rho = np.empty((n,), dtype=np.float64)
        status = np.empty((n,), dtype=np.uint8)
        n_ok = ctypes.c_int64(0)
        _lib.call("feos_host_pcsaft_liquid_density", _lib.ptr(p), _lib.ptr(t), _lib.ptr(pr), n, _lib.ptr(rho),
                  _lib.ptr(status), ctypes.cast(ctypes.byref(n_ok), ctypes.c_void_p))
        return rho[: n_ok.value].copy(), status.astype(bool)
