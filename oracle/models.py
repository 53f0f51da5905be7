"""TEST INFRASTRUCTURE ONLY -- oracle model classes with the reference's public API shape.

``PcSaftPure`` / ``PcSaftMix`` / ``GcPcSaftMix`` built from ``oracle.eos`` (restated Helmholtz
models + Newton-step formulas, autograd gradients) and ``oracle.solver`` (CPU solver).  Same
signatures, return orders, units and row-removal semantics as
`/root/reference/feos_torch/pcsaft_pure.py:184-233`, `pcsaft_mix.py:422-468`,
`gc_pcsaft.py:470-512` -- except that the model object is not mutated (`reduce`), which is not
observable through the returned values.
"""
import torch

from . import eos, solver


class PcSaftPure:
    """``compiled=True``: the solve stage runs in the compiled per-state solver (oracle/cpu_solver.cpp, OpenMP over
    states like the reference's rayon map, src/pcsaft.rs:86-92) instead of the batch-vectorised torch one; the two are
    the same algorithm and are checked against each other in tests/test_oracle.py."""

    def __init__(self, parameters, compiled=False):
        self.parameters = parameters
        if compiled:
            from . import cpu_solver

            self._solver = cpu_solver
        else:
            self._solver = solver

    def helmholtz_energy(self, temperature, density):
        return eos.pure_helmholtz(self.parameters, temperature, density)

    def derivatives(self, temperature, density):
        return eos.pure_derivatives(self.parameters, temperature, density)

    def vapor_pressure(self, temperature):
        rv, rl, nans = self._solver.pure_vle(self.parameters, temperature)
        ok = ~nans
        return nans, eos.pure_vapor_pressure_from_densities(self.parameters[ok], temperature[ok], rv[ok], rl[ok])

    def liquid_density(self, temperature, pressure):
        rho, nans = self._solver.pure_liquid_density(self.parameters, temperature, pressure)
        ok = ~nans
        return nans, eos.pure_liquid_density_from_density(self.parameters[ok], temperature[ok], pressure[ok], rho[ok])

    def equilibrium_liquid_density(self, temperature):
        rv, rl, nans = self._solver.pure_vle(self.parameters, temperature)
        ok = ~nans
        return nans, eos.pure_equilibrium_liquid_density_from_densities(self.parameters[ok], temperature[ok], rv[ok], rl[ok])


class _BinaryBase:
    def _helm(self, sel=None):
        raise NotImplementedError

    def _rmax(self):
        raise NotImplementedError

    def helmholtz_energy_density(self, temperature, density):
        return self._helm()(temperature, density)[:, None]

    def derivatives(self, temperature, density):
        return eos.mixture_derivatives(self._helm(), temperature, density)

    def _point(self, temperature, molefracs, pressure, bubble):
        rho, nans = solver.binary_bubble_dew(self._helm(), self._rmax(), temperature, molefracs, pressure, bubble)
        ok = ~nans
        helm = self._helm(ok)
        fn = eos.bubble_pressure_from_densities if bubble else eos.dew_pressure_from_densities
        return fn(helm, temperature[ok], rho[ok, 0:2], rho[ok, 2:4]), nans

    def bubble_point(self, temperature, liquid_molefracs, pressure):
        return self._point(temperature, liquid_molefracs, pressure, True)

    def dew_point(self, temperature, vapor_molefracs, pressure):
        return self._point(temperature, vapor_molefracs, pressure, False)


class PcSaftMix(_BinaryBase):
    def __init__(self, parameters, kij=None):
        self.parameters, self.kij = parameters, kij

    def _helm(self, sel=None):
        P, K = (self.parameters, self.kij) if sel is None else (self.parameters[sel], self.kij[sel])
        return lambda T, rho: eos.mix_helmholtz(P, K, T, rho)

    def _rmax(self):
        return solver.mix_max_density(self.parameters)


class GcPcSaftMix(_BinaryBase):
    def __init__(self, segment_identifier, parameter, segment_lists, bond_lists, binary_segment_records, phi=None):
        self.args = (segment_identifier, parameter, segment_lists, bond_lists, binary_segment_records)
        self.phi = phi
        self.tab = eos.GcTables(*self.args, phi)

    def _helm(self, sel=None):
        tab = self.tab
        if sel is not None:
            ident, par, segs, bonds, recs = self.args
            idx = torch.nonzero(sel).flatten().tolist()
            tab = eos.GcTables(ident, par, [segs[i] for i in idx], [bonds[i] for i in idx], recs,
                               None if self.phi is None else self.phi[sel])
        return lambda T, rho: eos.gc_helmholtz(tab, T, rho)

    def _rmax(self):
        return solver.gc_max_density(self.tab)
