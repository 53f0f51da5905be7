"""GPU parity tests of the differentiable state functions (helmholtz_energy / derivatives) through the
public API against torch autograd through the UNMODIFIED reference Python (tests/golden/state_grad.json,
``python -m oracle.make_golden state_grad``)."""
import numpy as np
import pytest
import torch

import cases

pytestmark = pytest.mark.gpu
F64 = torch.float64


@pytest.fixture(scope="module")
def ft():
    import __graft_entry__ as g

    g.build()
    import feos_torch_b200

    return feos_torch_b200


def close(got, want, tol):
    got = got.detach().cpu().numpy() if isinstance(got, torch.Tensor) else np.asarray(got)
    want = np.asarray(want, dtype=np.float64)
    return np.abs(got - want).max() <= tol * max(np.abs(want).max(), 1e-300)


def cu(x, grad=True):
    return torch.tensor(x, dtype=F64, device="cuda", requires_grad=grad)


def test_pure(ft, golden_state):
    e = golden_state["pure"]
    x, t, r, w = cu(e["params"]), cu(e["T"]), cu(e["rho"]), cu(e["w"], False)
    a, p, dp = ft.PcSaftPure(x).derivatives(t, r)
    assert close(a, e["a"], 1e-13) and close(p, e["p"], 1e-13) and close(dp, e["dp"], 1e-13)
    (w[:, 0] * a + w[:, 1] * p + w[:, 2] * dp).sum().backward()
    for i in range(len(e["T"])):
        assert close(x.grad[i], e["grad_params"][i], 1e-9), i
    assert close(t.grad, e["grad_T"], 1e-9) and close(r.grad, e["grad_rho"], 1e-9)
    # helmholtz_energy alone, host tensors in and out
    xh = torch.tensor(e["params"], dtype=F64, requires_grad=True)
    ah = ft.PcSaftPure(xh).helmholtz_energy(torch.tensor(e["T"], dtype=F64), torch.tensor(e["rho"], dtype=F64))
    assert not ah.is_cuda and close(ah, e["a"], 1e-13)
    ah.sum().backward()
    assert xh.grad is not None and torch.isfinite(xh.grad).all()


def test_mix(ft, golden_state):
    e = golden_state["mix"]
    x, k = cu(cases.MIX_PARAMS), cu(cases.MIX_KIJ)
    t, r, w = cu(e["T"]), cu(e["rho"]), cu(e["w"], False)
    a, p, mu, v = ft.PcSaftMix(x, k).derivatives(t, r)
    (w[:, 0] * a + w[:, 1] * p + (w[:, 2:4] * mu).sum(1) + (w[:, 4:6] * v).sum(1)).sum().backward()
    for i in range(len(e["T"])):
        label = cases.MIX_LABELS[i]
        assert close(x.grad[i].reshape(-1), np.array(e["grad_params"][i]).reshape(-1), 1e-8), label
        want_k = np.array(e["grad_kij"][i])
        assert np.abs(k.grad[i].cpu().numpy() - want_k).max() <= 1e-8 * max(np.abs(want_k).max(), 1e-300) + 1e-300, label
        assert close(t.grad[i], e["grad_T"][i], 1e-8) and close(r.grad[i], e["grad_rho"][i], 1e-8), label
    hd = ft.PcSaftMix(x.detach(), k.detach()).helmholtz_energy_density(t.detach(), r.detach())
    assert hd.shape == (len(e["T"]), 1) and torch.equal(hd[:, 0], a.detach())


def test_gc(ft, golden_state, golden_gc):
    e = golden_state["gc"]
    ident = list(golden_gc["segments"].keys())
    par = tuple(cu([golden_gc["segments"][i][j] for i in ident]) for j in range(8))
    kab = [cu(k[2]) for k in cases.GC_KAB]
    n = len(cases.GC_SEGMENTS)
    phi, t, r, w = cu([cases.GC_PHI] * n), cu(e["T"]), cu(e["rho"]), cu(e["w"], False)
    model = ft.GcPcSaftMix(ident, par, cases.GC_SEGMENTS, cases.GC_BONDS, [(a_, b_, kv) for (a_, b_, _), kv in zip(cases.GC_KAB, kab)], phi)
    a, p, mu, v = model.derivatives(t, r)
    (w[:, 0] * a + w[:, 1] * p + (w[:, 2:4] * mu).sum(1) + (w[:, 4:6] * v).sum(1)).sum().backward()
    for j in (0, 1, 3, 4, 5, 6, 7):  # eps_k, phi: NaN in the reference
        assert close(par[j].grad, e["grad_seg"][j], 1e-8), j
    assert close(torch.stack([k.grad for k in kab]), e["grad_kab"], 1e-8)
    assert close(t.grad, e["grad_T"], 1e-8) and close(r.grad, e["grad_rho"], 1e-8)
    assert torch.isfinite(phi.grad).all() and torch.isfinite(par[2].grad).all()
