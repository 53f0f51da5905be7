"""CPU tests of the host-side class hints (FEOS_PLAIN_ONLY): PcSaftMix._check_association on CPU tensors (it only
looks at the parameters) and GcPcSaftMix's _check_association on hand-made pair records."""
import numpy as np
import pytest
import torch

from feos_torch_b200 import gc_pcsaft, pcsaft_mix

F64 = torch.float64


def _pair(mu=0.0, sites=(0.0, 0.0)):
    a = [1.5, 3.4, 180.0, mu, 0.02 if any(sites) else 0.0, 2000.0 if any(sites) else 0.0, sites[0], sites[1]]
    b = [2.0, 3.6, 230.0, 0.0, 0.0, 0.0, 0.0, 0.0]
    return [a, b]


def test_mix_plain_hint_and_cache():
    check = pcsaft_mix.PcSaftMix._check_association
    pcsaft_mix._CLASS_CACHE.clear()
    plain = torch.tensor([_pair()] * 3, dtype=F64)
    assert check(None, plain) is True
    assert len(pcsaft_mix._CLASS_CACHE) == 1 and check(None, plain) is True and len(pcsaft_mix._CLASS_CACHE) == 1
    assert check(None, torch.tensor([_pair(), _pair(mu=1.8)], dtype=F64)) is False       # a polar pair in the batch
    assert check(None, torch.tensor([_pair(sites=(1.0, 1.0))], dtype=F64)) is False      # association sites
    # one row expanded over the batch is recognised through its base; an in-place change of the row is noticed
    row = torch.tensor(_pair(), dtype=F64)
    assert check(None, row.expand(5, 2, 8)) is True
    with torch.no_grad():
        row[0, 3] = 2.0
    assert check(None, row.expand(5, 2, 8)) is False
    assert len(pcsaft_mix._CLASS_CACHE) <= 8


def test_mix_too_many_associating_components():
    three = torch.tensor([[_pair(sites=(1.0, 1.0))[0]] * 3], dtype=F64)  # [1, 3, 8]: three associating components
    with pytest.raises(Exception, match="two associating components"):
        pcsaft_mix.PcSaftMix._check_association(None, three)


def _rec(nseg, seg, cnt0, cnt1, nbond):
    rec = {"nseg": nseg, "nbond": nbond, "seg": np.array(seg + [0] * (12 - len(seg))),
           "cnt": np.zeros((2, 12))}
    rec["cnt"][0, :nseg] = cnt0
    rec["cnt"][1, :nseg] = cnt1
    return rec


def test_gc_slim_hint():
    # segment table rows: m, sigma, eps, mu, kappa_ab, eps_ab, na, nb -- row 2 associates
    seg = torch.tensor([[0.7, 3.7, 230.0, 0, 0, 0, 0, 0], [0.6, 3.9, 260.0, 2.0, 0, 0, 0, 0],
                        [0.4, 3.0, 350.0, 0, 0.01, 2500.0, 1, 1]], dtype=F64)
    plain = _rec(2, [0, 1], [2, 0], [1, 1], 2)
    assoc = _rec(2, [0, 2], [2, 0], [1, 1], 2)
    many_bonds = _rec(2, [0, 1], [2, 0], [1, 1], gc_pcsaft.SLIM_MAX_BONDS + 1)
    assert gc_pcsaft._check_association([plain], seg) is True
    assert gc_pcsaft._check_association([plain, assoc], seg) is False
    assert gc_pcsaft._check_association([many_bonds], seg) is False
    two = _rec(2, [0, 2], [0, 2], [1, 1], 2)  # two associating segments in molecule 0
    with pytest.raises(Exception, match="one associating segment"):
        gc_pcsaft._check_association([two], seg)
