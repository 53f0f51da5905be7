"""Limits of the REFERENCE that the scope of this package rests on, demonstrated by running the unmodified reference
Python (build container only).  SURVEY.md section 8(f) lists "n > 2 non-associating mixtures" as something the
reference's Helmholtz energy supports; through its public API it does not."""
import pytest
import torch

from oracle import refshim

pytestmark = pytest.mark.skipif(not refshim.reference_available(), reason="/root/reference not present (GPU box)")
F64 = torch.float64


def test_reference_cannot_evaluate_three_components():
    ref = refshim.load_reference()
    P = torch.tensor([[[1.5, 3.2, 150, 0, 0, 0, 0, 0], [2.5, 3.5, 250, 0, 0, 0, 0, 0], [2.0, 3.4, 200, 0, 0, 0, 0, 0]]], dtype=F64)
    T = torch.tensor([300.0], dtype=F64)
    rho = torch.tensor([[0.001, 0.002, 0.0015]], dtype=F64)
    with pytest.raises(Exception, match="binary"):  # pcsaft_mix.py:74-75
        ref.PcSaftMix(P, torch.zeros((1, 2), dtype=F64)).helmholtz_energy_density(T, rho)
    with pytest.raises(TypeError):  # pcsaft_mix.py:141: self.kij[cross_associating, 1] with kij = None
        ref.PcSaftMix(P, None).helmholtz_energy_density(T, rho)


def test_this_package_rejects_it_the_same_way():
    from feos_torch_b200 import PcSaftMix

    with pytest.raises((NotImplementedError, ValueError)):
        PcSaftMix(torch.zeros((1, 3, 8), dtype=F64), None)


def test_reference_kab_argument_is_dead():
    """GcPcSaftMix.helmholtz_energy_density(T, rho, kab=...) never reads `kab` (gc_pcsaft.py:116 declares it, :188
    uses self.kab): passing a different value changes nothing."""
    import json
    import os

    import cases

    ref = refshim.load_reference()
    table = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "gc.json")))["segments"]
    ident = list(table.keys())
    par = tuple(torch.tensor([table[i][k] for i in ident], dtype=F64) for k in range(8))
    phi = torch.tensor([cases.GC_PHI], dtype=F64)
    m = ref.GcPcSaftMix(ident, par, [cases.GC_SEGMENTS[2]], [cases.GC_BONDS[2]], cases.GC_KAB, phi)
    T = torch.tensor([300.0], dtype=F64)
    rho = torch.tensor([[0.001, 0.002]], dtype=F64)
    a = m.helmholtz_energy_density(T, rho)
    b = m.helmholtz_energy_density(T, rho, kab=torch.full((len(ident), len(ident)), 0.5, dtype=F64))
    assert torch.equal(a, b)
