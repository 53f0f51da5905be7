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

    @staticmethod
    def _point(name, parameters, kij, temperature, molefracs, pressure):
        p, k, t, z, pr = _f64(parameters), _f64(kij), _f64(temperature), _f64(molefracs), _f64(pressure)
        n = t.shape[0]
        if p.shape != (n, 2, 8) or k.shape != (n, 2) or z.shape != (n,):
            raise ValueError("parameters must have shape [n, 2, 8], kij [n, 2], molefracs [n]")
        rho = np.empty((n, 4), dtype=np.float64)
        status = np.empty((n,), dtype=np.uint8)
        n_ok = ctypes.c_int64(0)
        _lib.call(name, _lib.ptr(p), _lib.ptr(k), _lib.ptr(t), _lib.ptr(z), _lib.ptr(pr), n, _lib.ptr(rho),
                  _lib.ptr(status), ctypes.cast(ctypes.byref(n_ok), ctypes.c_void_p))
        return rho[: n_ok.value].copy(), status.astype(bool)

    @staticmethod
    def bubble_point(parameters, kij, temperature, liquid_molefracs, pressure):
        """-> (rho f64[n_ok,4] = (rho_V1, rho_V2, rho_L1, rho_L2) in 1/A^3, failed bool[n]); src/pcsaft.rs:44-60."""
        return PcSaft._point("feos_host_pcsaft_bubble_point", parameters, kij, temperature, liquid_molefracs, pressure)

    @staticmethod
    def dew_point(parameters, kij, temperature, vapor_molefracs, pressure):
        """-> same tuple as bubble_point; src/pcsaft.rs:63-79."""
        return PcSaft._point("feos_host_pcsaft_dew_point", parameters, kij, temperature, vapor_molefracs, pressure)
