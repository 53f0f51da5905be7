"""CPU tests of the gradients of the public state functions (helmholtz_energy / derivatives) as the
kernels compute them (csrc/state_grad.cuh compiled for the host) against torch autograd through the
UNMODIFIED reference Python (golden file tests/golden/state_grad.json, made by
``python -m oracle.make_golden state_grad``; pcsaft_pure.py:180-182, pcsaft_mix.py:395-420,
gc_pcsaft.py:443-468) on the parameter sets of the reference's own tests (tests/cases.py)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "hostsim"))
import sim  # noqa: E402

import cases
from feos_torch_b200.gc_pcsaft import build_pair_table


def close(got, want, tol=1e-9):
    got, want = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
    return np.abs(got - want).max() <= tol * max(np.abs(want).max(), 1e-300)


def test_pure_state_gradients(golden_state):
    e = golden_state["pure"]
    w = np.array(e["w"])
    g = sim.pure_state_bwd(np.array(e["params"]), np.array(e["T"]), np.array(e["rho"]), w[:, 0], w[:, 1], w[:, 2])
    for i in range(len(e["T"])):
        assert close(g[i, :8], e["grad_params"][i]), i
    assert close(g[:, 8], e["grad_T"]) and close(g[:, 9], e["grad_rho"])


def test_mix_state_gradients(golden_state):
    e = golden_state["mix"]
    P, K = np.array(cases.MIX_PARAMS, dtype=np.float64), np.array(cases.MIX_KIJ)
    g = sim.mix_state_bwd(P, K, np.array(e["T"]), np.array(e["rho"]), np.array(e["w"]))
    for i in range(P.shape[0]):  # per row: the magnitudes differ by orders between categories
        label = cases.MIX_LABELS[i]
        assert close(g[i, :16], np.array(e["grad_params"][i]).reshape(-1), 1e-8), label
        want_k = np.array(e["grad_kij"][i])
        assert np.abs(g[i, 16:18] - want_k).max() <= 1e-8 * max(np.abs(want_k).max(), 1e-300) + 1e-300, label
        assert close(g[i, 18], e["grad_T"][i], 1e-8) and close(g[i, 19:21], e["grad_rho"][i], 1e-8), label


def test_gc_state_gradients(golden_state, golden_gc):
    e = golden_state["gc"]
    ident = list(golden_gc["segments"].keys())
    seg = np.array([golden_gc["segments"][i] for i in ident])
    n = len(cases.GC_SEGMENTS)
    pairs, idx = build_pair_table(ident, cases.GC_SEGMENTS, cases.GC_BONDS, cases.GC_KAB)
    gs, gk, gp, gt, gr = sim.gc_state_bwd(pairs, seg, np.array([k[2] for k in cases.GC_KAB]), idx, np.array([cases.GC_PHI] * n),
                                          np.array(e["T"]), np.array(e["rho"]), np.array(e["w"]))
    for j in (0, 1, 3, 4, 5, 6, 7):  # eps_k and phi: NaN in the reference (sqrt(0) of the eps = 0 segment)
        assert close(gs[:, j], e["grad_seg"][j], 1e-8), j
    assert close(gk, e["grad_kab"], 1e-8)
    assert close(gt, e["grad_T"], 1e-8) and close(gr, e["grad_rho"], 1e-8)
    assert np.isfinite(gp).all() and np.isfinite(gs).all()
