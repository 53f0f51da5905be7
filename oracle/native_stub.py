"""TEST INFRASTRUCTURE ONLY.  numpy-in / numpy-out stand-ins for the reference's native classes
``PcSaft`` (`/root/reference/src/pcsaft.rs:13-80`) and ``GcPcSaft`` (`src/gc_pcsaft.rs:15-100`),
backed by ``oracle.solver``.  Same tuples: compacted densities of the converged states and a
bool mask of length N with True = failed."""
import numpy as np
import torch

from . import eos, solver


def _t(x):
    return torch.from_numpy(np.ascontiguousarray(x, dtype=np.float64))


class PcSaft:
    @staticmethod
    def vapor_pressure(parameters, temperature):
        rv, rl, failed = solver.pure_vle(_t(parameters), _t(temperature))
        ok = ~failed
        rho = torch.zeros((int(ok.sum()), 4), dtype=torch.float64)  # cols 2,3 stay 0 (pcsaft.rs:95-101)
        rho[:, 0] = rv[ok]
        rho[:, 1] = rl[ok]
        return rho.numpy(), failed.numpy()

    @staticmethod
    def liquid_density(parameters, temperature, pressure):
        rho, failed = solver.pure_liquid_density(_t(parameters), _t(temperature), _t(pressure))
        return rho[~failed].numpy(), failed.numpy()

    @staticmethod
    def _bubble_dew(parameters, kij, temperature, molefracs, pressure, bubble):
        P, K = _t(parameters), _t(kij)
        helm = lambda T, rho: eos.mix_helmholtz(P, K, T, rho)
        rho, failed = solver.binary_bubble_dew(helm, solver.mix_max_density(P), _t(temperature), _t(molefracs), _t(pressure), bubble)
        return rho[~failed].numpy(), failed.numpy()

    @staticmethod
    def bubble_point(parameters, kij, temperature, liquid_molefracs, pressure):
        return PcSaft._bubble_dew(parameters, kij, temperature, liquid_molefracs, pressure, True)

    @staticmethod
    def dew_point(parameters, kij, temperature, vapor_molefracs, pressure):
        return PcSaft._bubble_dew(parameters, kij, temperature, vapor_molefracs, pressure, False)


class GcPcSaft:
    def __init__(self, segment_records, segments, bonds, binary_segment_records, phi):
        ident = [r[0] for r in segment_records]
        cols = np.stack([np.asarray(r[1], dtype=np.float64) for r in segment_records], axis=1)  # [8,S]
        parameter = tuple(torch.from_numpy(np.ascontiguousarray(c)) for c in cols)
        recs = [(s1, s2, float(k)) for s1, s2, k in binary_segment_records]
        self.tab = eos.GcTables(ident, parameter, segments, bonds, recs, _t(phi))

    def _bubble_dew(self, temperature, molefracs, pressure, bubble):
        helm = lambda T, rho: eos.gc_helmholtz(self.tab, T, rho)
        rho, failed = solver.binary_bubble_dew(helm, solver.gc_max_density(self.tab), _t(temperature), _t(molefracs), _t(pressure), bubble)
        return rho[~failed].numpy(), failed.numpy()

    def bubble_point(self, temperature, liquid_molefracs, pressure):
        return self._bubble_dew(temperature, liquid_molefracs, pressure, True)

    def dew_point(self, temperature, vapor_molefracs, pressure):
        return self._bubble_dew(temperature, vapor_molefracs, pressure, False)
