"""CPU tests of the group-contribution device functions (coefficient stage, shared mixture model,
solver, gradient passes; compiled for the host, tests/hostsim) and of the pair-table builder against
the golden vectors of the reference Python (tests/test_gc_pcsaft.py inputs)."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "hostsim"))
import sim  # noqa: E402

import cases
from feos_torch_b200.gc_pcsaft import PAIR_DTYPE, build_pair_table


def table(golden_gc):
    ident = list(golden_gc["segments"].keys())
    return ident, np.array([golden_gc["segments"][i] for i in ident])


def test_pair_table_builder(golden_gc):
    ident, _ = table(golden_gc)
    segs = [cases.GC_SEGMENTS[0], cases.GC_SEGMENTS[1], cases.GC_SEGMENTS[0]]
    bonds = [cases.GC_BONDS[0], cases.GC_BONDS[1], cases.GC_BONDS[0]]
    pairs, idx = build_pair_table(ident, segs, bonds, [("CH3", "CH2", -0.15)])
    assert PAIR_DTYPE.itemsize == 776 and idx.tolist() == [0, 1, 0] and len(pairs) == 2
    p = pairs[0]  # butane / propane: CH3, CH2
    assert p["nseg"] == 2 and [ident[i] for i in p["seg"][:2]] == ["CH3", "CH2"]
    assert p["cnt"][:, :2].tolist() == [[2.0, 2.0], [2.0, 1.0]]
    bt = {(int(a), int(b)): (p["bcnt"][0, k], p["bcnt"][1, k]) for k, (a, b) in enumerate(zip(p["ba"][: p["nbond"]], p["bb"][: p["nbond"]]))}
    assert bt == {(1, 0): (2.0, 2.0), (1, 1): (1.0, 0.0)}  # CH2-CH3 twice each; CH2-CH2 once in butane
    assert p["krec"][0, 1] == 0 and p["krec"][1, 0] == 0 and p["krec"][0, 0] == -1


def test_gc_state_functions(golden_gc):
    ident, seg = table(golden_gc)
    e = golden_gc["test_grid"]
    n = len(e["T"])
    pairs, idx = build_pair_table(ident, cases.GC_SEGMENTS, cases.GC_BONDS, cases.GC_KAB)
    kab = np.array([k[2] for k in cases.GC_KAB])
    phi = np.array([cases.GC_PHI] * n)
    a, p, mu, v = sim.gc_state(pairs, seg, kab, idx, phi, np.array(e["T"]), np.array(e["rho"]))
    # tests/test_gc_pcsaft.py:122-127: abs 1e-14 on a, mu, p; 1e-11 on v
    assert np.abs(a - e["a"]).max() < 1e-14
    assert np.abs(p - e["p_red"]).max() < 1e-14
    assert np.abs(mu - e["mu"]).max() < 1e-13
    assert (np.abs(v - e["v"]) / np.abs(e["v"])).max() < 1e-11


@pytest.mark.parametrize("key,bubble", [("bubble", True), ("dew", False)])
def test_gc_points(golden_gc, key, bubble):
    ident, seg = table(golden_gc)
    e = golden_gc["category_grid"]
    rows = e["category"]
    pairs, idx = build_pair_table(ident, [cases.GC_SEGMENTS[i] for i in rows], [cases.GC_BONDS[i] for i in rows], cases.GC_KAB)
    assert len(pairs) == 11
    kab = np.array([k[2] for k in cases.GC_KAB])
    T, z = np.array(e["T"]), np.array(e["z"])
    n = len(T)
    phi = np.array([cases.GC_PHI] * n)
    ge = e[key]
    rho, val, st, ev = sim.gc_point(pairs, seg, kab, idx, phi, T, z, bubble)
    assert (st != 0).tolist() == ge["failed"]
    assert (np.abs(val - ge["value"]) / np.abs(ge["value"])).max() < 1e-10
    gs, gk, gp, gt = sim.gc_point_bwd(pairs, seg, kab, idx, phi, T, rho, bubble, np.ones(n))
    want = np.stack([np.array(ge["grad_seg%d" % j]) for j in range(8)], axis=1)
    for col in (0, 1, 3, 4, 5, 6, 7):
        assert np.abs(gs[:, col] - want[:, col]).max() <= 1e-8 * np.abs(want[:, col]).max() + 1e-300
    # the reference's epsilon_k and phi gradients are NaN for this table (sqrt(0) of the eps = 0 segment
    # ">C<"): check ours against central differences of the solver output
    assert np.isnan(want[:, 2]).all() and np.isnan(np.array(ge["grad_phi"])).any()
    assert np.isfinite(gs).all() and np.isfinite(gp).all()
    j = ident.index("CH3")
    h = 1e-5 * seg[j, 2]
    sp, sm = seg.copy(), seg.copy()
    sp[j, 2] += h
    sm[j, 2] -= h
    fd = (sim.gc_point(pairs, sp, kab, idx, phi, T, z, bubble)[1] - sim.gc_point(pairs, sm, kab, idx, phi, T, z, bubble)[1]).sum() / (2 * h)
    assert abs(fd - gs[j, 2]) < 1e-6 * abs(fd)
    pp, pm = phi.copy(), phi.copy()
    pp[:, 0] += 1e-6
    pm[:, 0] -= 1e-6
    fd = (sim.gc_point(pairs, seg, kab, idx, pp, T, z, bubble)[1] - sim.gc_point(pairs, seg, kab, idx, pm, T, z, bubble)[1]) / 2e-6
    assert (np.abs(fd - gp[:, 0]) <= 1e-6 * np.abs(fd) + 1e-8 * np.abs(val)).all()  # FD noise floor ~1e-8 |p|
    assert np.abs(gk - [ge["grad_kab0"], ge["grad_kab1"]]).max() < 1e-8 * np.abs(gk).max()
    assert (np.abs(gt - ge["grad_T"]) / np.abs(ge["grad_T"])).max() < 1e-8


@pytest.mark.parametrize("bubble", [True, False])
def test_slim_model_matches_general(golden_gc, bubble):
    """The slim-class instantiation of the GC forward kernel (no association term, <= 8 bond types: 45 doubles per
    state, bond counts read from the pair table) runs the same solver on the same numbers: wherever its fast path
    finishes, densities, pressure and evaluation count equal the general model's bit for bit; the associating
    categories are left to the general instantiation."""
    ident, seg = table(golden_gc)
    e = golden_gc["category_grid"]
    rows = e["category"]
    pairs, idx = build_pair_table(ident, [cases.GC_SEGMENTS[i] for i in rows], [cases.GC_BONDS[i] for i in rows], cases.GC_KAB)
    kab = np.array([k[2] for k in cases.GC_KAB])
    T, z = np.array(e["T"]), np.array(e["z"])
    phi = np.array([cases.GC_PHI] * len(T))
    rho_s, val_s, st_s, ev_s = sim.gc_point_slim(pairs, seg, kab, idx, phi, T, z, bubble)
    rho, val, st, ev = sim.gc_point(pairs, seg, kab, idx, phi, T, z, bubble)
    slim = st_s != 3
    assert 8 <= slim.sum() < len(T)  # both classes occur on the grid
    done = slim & (st_s == 0)
    assert done.sum() >= 8 and (st[done] == 0).all()
    # (the dipole term is evaluated from per-state polynomial coefficients in the slim model: same value to round-off)
    assert (np.abs(rho_s[done] - rho[done]) / np.abs(rho[done])).max() < 1e-12
    assert (np.abs(val_s[done] - val[done]) / np.abs(val[done])).max() < 1e-12
    assert np.array_equal(ev_s[done], ev[done])


def test_slim_evaluation_matches_reference_state_functions(golden_gc):
    """a, p, mu, v of the slim model (dipole term from DipolePoly) against the reference Python at the golden
    (T, rho) states, and its Hessian against the general closed-form evaluation."""
    ident, seg = table(golden_gc)
    e = golden_gc["test_grid"]
    pairs, idx = build_pair_table(ident, cases.GC_SEGMENTS, cases.GC_BONDS, cases.GC_KAB)
    kab = np.array([k[2] for k in cases.GC_KAB])
    T, rho = np.array(e["T"]), np.array(e["rho"])
    phi = np.array([cases.GC_PHI] * len(T))
    a, p, mu, v, h, slim = sim.gc_state_slim(pairs, seg, kab, idx, phi, T, rho)
    assert slim.any() and not slim.all()
    assert np.abs(a - e["a"]).max() < 1e-14 and np.abs(p - e["p_red"]).max() < 1e-14
    assert np.abs(mu - e["mu"]).max() < 1e-13
    assert (np.abs(v - e["v"]) / np.abs(e["v"])).max() < 1e-11


def test_reference_gc_configs(golden_gc):
    """tests/test_gc_pcsaft.py:130-222: butane/propane, k_ab(CH3, CH2) = -0.15, T = 150 K, z = 0.5."""
    ident, seg = table(golden_gc)
    cfg = cases.GC_POINT
    pairs, idx = build_pair_table(ident, [cfg["segments"]], [cfg["bonds"]], cfg["kab"])
    kab = np.array([cfg["kab"][0][2]])
    phi = np.array([cfg["phi"]])
    for name, bubble in (("test_bubble_point", True), ("test_dew_point", False)):
        g = golden_gc[name]
        rho, val, st, ev = sim.gc_point(pairs, seg, kab, idx, phi, np.array([cfg["T"]]), np.array([cfg["z"]]), bubble)
        assert st[0] == 0 and abs(val[0] - g["value"][0]) < 1e-8  # abs 1e-8 Pa, tests/test_gc_pcsaft.py:173
        gs, gk, gp, gt = sim.gc_point_bwd(pairs, seg, kab, idx, phi, np.array([cfg["T"]]), rho, bubble, np.ones(1))
        assert abs(gk[0] - g["grad_kab"][0]) < 1e-8 * abs(g["grad_kab"][0])
        assert abs(gt[0] - g["grad_T"][0]) < 1e-8 * abs(g["grad_T"][0])


def test_base_adjoint_equals_dual_passes(golden_gc):
    """The reverse-mode sweep over the base coefficient group (+ dual passes restricted to the dipole /
    association groups) against dual passes over every input direction and every coefficient."""
    ident, seg = table(golden_gc)
    e = golden_gc["category_grid"]
    rows = e["category"]
    pairs, idx = build_pair_table(ident, [cases.GC_SEGMENTS[i] for i in rows], [cases.GC_BONDS[i] for i in rows], cases.GC_KAB)
    kab = np.array([k[2] for k in cases.GC_KAB])
    T, z = np.array(e["T"]), np.array(e["z"])
    n = len(T)
    phi = np.array([cases.GC_PHI] * n)
    g = np.linspace(0.5, 1.5, n)
    for bubble in (True, False):
        rho = sim.gc_point(pairs, seg, kab, idx, phi, T, z, bubble)[0]
        a = sim.gc_point_bwd(pairs, seg, kab, idx, phi, T, rho, bubble, g, generic=True)
        b = sim.gc_point_bwd(pairs, seg, kab, idx, phi, T, rho, bubble, g)
        for x, y in zip(a, b):
            assert np.abs(x - y).max() <= 1e-10 * np.abs(x).max()
        # the k_ab-only path (gc_point_bwd_kab_kernel: two closed-form adjoints + the record pairs) gives the same k_ab gradient
        gk = sim.gc_point_bwd_kab(pairs, seg, kab, idx, phi, T, rho, bubble, g)
        assert np.abs(gk - b[1]).max() <= 1e-12 * np.abs(b[1]).max() and np.abs(b[1]).min() > 0


def test_closed_form_base_eval_equals_hyper_dual(golden_gc):
    ident, seg = table(golden_gc)
    n = len(cases.GC_SEGMENTS)
    pairs, idx = build_pair_table(ident, cases.GC_SEGMENTS, cases.GC_BONDS, cases.GC_KAB)
    kab = np.array([k[2] for k in cases.GC_KAB])
    rng = np.random.default_rng(3)
    reps = 50
    idx = np.tile(idx, reps)
    T = rng.uniform(200, 400, n * reps)
    rho = np.exp(rng.uniform(np.log(1e-8), np.log(3e-3), (n * reps, 2)))
    phi = np.array([cases.GC_PHI] * (n * reps))
    fast, generic = sim.gc_base_eval(pairs, seg, kab, idx, phi, T, rho)
    err = np.abs(fast - generic) / np.maximum(np.abs(generic), 1e-300)
    assert err.max() < 1e-10 and np.median(err) < 1e-14
