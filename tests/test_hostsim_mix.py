"""CPU tests of the binary-mixture device functions (Helmholtz model, bubble/dew solver, gradient
passes; compiled for the host, tests/hostsim) against the golden vectors of the reference Python."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "hostsim"))
import sim  # noqa: E402

# random grid only: next to a critical point (liquid / vapour density ratio below this) converged-or-failed is
# implementation-defined, like tau in (0.97, 1.03) for pure components
NEAR_CRITICAL_RATIO = 1.5


def test_mix_state_functions(golden_mix):
    e = golden_mix["test_grid"]
    P = np.array(e["params"]).reshape(-1, 16)
    a, p, mu, v = sim.mix_state(P, np.array(e["kij"]), np.array(e["T"]), np.array(e["rho"]))
    # tests/test_pcsaft_mix.py:119-124: abs 1e-14 on a, mu, p and 1e-11 on v
    assert np.abs(a - e["a"]).max() < 1e-14
    assert np.abs(p - e["p_red"]).max() < 1e-14
    assert np.abs(mu - e["mu"]).max() < 1e-13
    assert (np.abs(v - e["v"]) / np.abs(e["v"])).max() < 1e-11


@pytest.mark.parametrize("grid", ["category_grid", "random_grid"])
@pytest.mark.parametrize("key,bubble", [("bubble", True), ("dew", False)])
def test_mix_points(golden_mix, key, bubble, grid):
    e = golden_mix[grid]
    P = np.array(e["params"]).reshape(-1, 16)
    K, T, z = np.array(e["kij"]), np.array(e["T"]), np.array(e["z"])
    ge = e[key]
    rho, val, st, ev = sim.mix_point(P, K, T, z, bubble)
    failed, gf = st != 0, np.array(ge["failed"])
    if grid == "category_grid":
        # all 14 categories x 3 temperatures x 3 compositions: identical mask, no exemptions
        assert failed.tolist() == gf.tolist() and len(gf) == 126
    else:
        ratio = np.array([r if r is not None else np.nan for r in ge["rho_ratio"]], dtype=float)
        mine = rho[:, 2:4].sum(axis=1) / rho[:, 0:2].sum(axis=1)
        with np.errstate(invalid="ignore"):
            near = (~gf & (ratio < NEAR_CRITICAL_RATIO)) | (~failed & (mine < NEAR_CRITICAL_RATIO))
        strict = ~((failed != gf) & near)
        assert failed[strict].tolist() == gf[strict].tolist() and strict.sum() >= 130
    both = ~failed & ~gf
    gv = np.full(len(T), np.nan)
    gv[~gf] = ge["value"]
    assert (np.abs(val[both] - gv[both]) / np.abs(gv[both])).max() < 1e-10
    grad = sim.mix_point_bwd(P[both], K[both], T[both], rho[both], bubble, np.ones(int(both.sum())))
    gp = np.array(ge["grad_params"]).reshape(-1, 16)[both]
    gk = np.array(ge["grad_kij"])[both]
    gT = np.array(ge["grad_T"])[both]
    scale = np.abs(gp).max(axis=1, keepdims=True)
    assert (np.abs(grad[:, :16] - gp) / scale).max() < 1e-8
    assert (np.abs(grad[:, 16:18] - gk) / (np.abs(gk).max(axis=1, keepdims=True) + 1e-300)).max() < 1e-8
    assert (np.abs(grad[:, 18] - gT) / np.abs(gT)).max() < 1e-8


@pytest.mark.parametrize("key,bubble", [("bubble", True), ("dew", False)])
def test_mix_kij_only_gradient(golden_mix, key, bubble):
    """mix_point_bwd_kij_kernel's closed form (k_ij enters through q1[01], q2[01] alone) against the reference
    autograd d p / d k_ij on all 14 categories -- polar and associating pairs included."""
    e = golden_mix["category_grid"]
    P = np.array(e["params"]).reshape(-1, 16)
    K, T, z = np.array(e["kij"]), np.array(e["T"]), np.array(e["z"])
    ge = e[key]
    rho, val, st, ev = sim.mix_point(P, K, T, z, bubble)
    ok = st == 0
    gk = np.array(ge["grad_kij"])[ok][:, 0]
    mine = sim.mix_point_bwd_kij(P[ok], K[ok], T[ok], rho[ok], bubble)
    assert (np.abs(mine - gk) / np.abs(gk)).max() < 1e-8


@pytest.mark.parametrize("bubble", [True, False])
def test_mix_plain_model_matches_general(golden_mix, bubble):
    """The plain-pair instantiation of the forward kernel runs the same solver on the same numbers: wherever it
    finishes (status 0) densities, pressure and evaluation count equal the general model's, bit for bit."""
    e = golden_mix["random_grid"]
    P = np.array(e["params"]).reshape(-1, 16)
    K, T, z = np.array(e["kij"]), np.array(e["T"]), np.array(e["z"])
    rho_p, val_p, st_p, ev_p = sim.mix_point_plain(P, K, T, z, bubble)
    rho, val, st, ev = sim.mix_point(P, K, T, z, bubble, robust=False)
    plain = st_p != 3
    assert plain.sum() >= 10 and (st_p[plain] == st[plain]).all()
    done = plain & (st_p == 0)
    assert np.array_equal(rho_p[done], rho[done]) and np.array_equal(val_p[done], val[done])
    assert np.array_equal(ev_p[done], ev[done])


def test_reference_bubble_dew_configs(golden_mix):
    """tests/test_pcsaft_mix.py:127-251: p and dp/dk_ij of the reference's own test configurations."""
    for name, bubble in (("test_bubble_point", True), ("test_dew_point", False)):
        g = golden_mix[name]
        P = np.array(g["params"]).reshape(-1, 16)
        K, T, z = np.array(g["kij"]), np.array(g["T"]), np.array(g["z"])
        rho, val, st, ev = sim.mix_point(P, K, T, z, bubble)
        assert st[0] == 0
        assert abs(val[0] - g["value"][0]) < 1e-8  # abs 1e-8 Pa, tests/test_pcsaft_mix.py:190
        grad = sim.mix_point_bwd(P, K, T, rho, bubble, np.ones(1))
        gk = np.array(g["grad_kij"])[0]
        assert np.abs(grad[0, 16:18] - gk).max() < 1e-8 * np.abs(gk).max()


def test_mix_c4_grid_converges():
    """BASELINE config C4 shape: shared parameter pair, k_ij in [-0.15, -0.05], T in [0.6, 0.85] min T_c."""
    rng = np.random.default_rng(3)
    n = 4000
    pa, pb = [1, 3.5, 150, 0, 0, 0, 0, 0], [1, 3.5, 200, 0, 0, 0, 0, 0]
    P = np.tile(np.array(pa + pb, dtype=float), (n, 1))
    K = np.zeros((n, 2))
    K[:, 0] = rng.uniform(-0.15, -0.05, n)
    T = rng.uniform(0.6, 0.85, n) * 150 * 1.2757
    z = rng.uniform(0.05, 0.95, n)
    for bubble in (True, False):
        rho, val, st, ev = sim.mix_point(P, K, T, z, bubble)
        assert (st == 0).all()
        _, pV, muV, _ = sim.mix_state(P, K, T, rho[:, 0:2])
        _, pL, muL, _ = sim.mix_state(P, K, T, rho[:, 2:4])
        assert (np.abs(pV - pL) / rho[:, 2:4].sum(axis=1)).max() < 1e-12
        assert np.abs(muV + np.log(rho[:, 0:2]) - muL - np.log(rho[:, 2:4])).max() < 1e-9
        comp = rho[:, 2] / rho[:, 2:4].sum(axis=1) if bubble else rho[:, 0] / rho[:, 0:2].sum(axis=1)
        assert np.abs(comp - z).max() < 1e-12


def test_closed_form_base_eval_equals_hyper_dual():
    """mix_base_eval (hand-derived value / gradient / Hessian of hard sphere + hard chain + dispersion) against
    the same terms in generic H2 arithmetic, from ideal-gas to liquid packing fractions."""
    rng = np.random.default_rng(0)
    n = 4000
    P = np.zeros((n, 2, 8))
    P[:, :, 0] = rng.uniform(1, 4, (n, 2))
    P[:, :, 1] = rng.uniform(3, 4, (n, 2))
    P[:, :, 2] = rng.uniform(150, 350, (n, 2))
    K = np.zeros((n, 2))
    K[:, 0] = rng.uniform(-0.1, 0.1, n)
    T = rng.uniform(150, 400, n)
    eta = np.exp(rng.uniform(np.log(1e-7), np.log(0.5), n))
    x = rng.uniform(0.01, 0.99, n)
    d = P[:, :, 1] * (1 - 0.12 * np.exp(-3 * P[:, :, 2] / T[:, None]))
    rt = eta / (np.pi / 6 * (x * P[:, 0, 0] * d[:, 0] ** 3 + (1 - x) * P[:, 1, 0] * d[:, 1] ** 3))
    rho = np.stack([rt * x, rt * (1 - x)], 1)
    fast, generic = sim.mix_base_eval(P.reshape(n, 16), K, T, rho)
    err = np.abs(fast - generic) / np.maximum(np.abs(generic), 1e-300)
    # both share the reference's cancellation at low density (zeta_2^3 / eta terms of the hard-sphere formula)
    assert err.max() < 1e-10 and np.median(err) < 1e-14


def test_point_machine_equals_loop_solver(golden_mix):
    """The resumable PointMachine the work-queue kernels run (one evaluation per trip) against the loop solver:
    same densities, values, status and evaluation counts on the category grid and on a C4-shaped random grid
    that includes states the fast path gives up on."""
    rng = np.random.default_rng(5)
    n = 3000
    pa, pb = [1, 3.5, 150, 0, 0, 0, 0, 0], [1, 3.5, 200, 0, 0, 0, 0, 0]
    P = np.tile(np.array(pa + pb, dtype=float), (n, 1))
    K = np.zeros((n, 2))
    K[:, 0] = rng.uniform(-0.15, -0.05, n)
    T = rng.uniform(0.6, 1.05, n) * 150 * 1.2757  # up to above the critical temperature of the light component
    z = rng.uniform(0.02, 0.98, n)
    e = golden_mix["category_grid"]
    grids = [(P, K, T, z), (np.array(e["params"]).reshape(-1, 16), np.array(e["kij"]), np.array(e["T"]), np.array(e["z"]))]
    try:
        for (Pg, Kg, Tg, zg) in grids:
            for bubble in (True, False):
                sim.use_point_machine(False)
                a = sim.mix_point(Pg, Kg, Tg, zg, bubble, robust=False)
                sim.use_point_machine(True)
                b = sim.mix_point(Pg, Kg, Tg, zg, bubble, robust=False)
                assert (a[2] == b[2]).all() and (a[3] == b[3]).all()
                ok = a[2] == 0
                assert ok.any() and np.array_equal(a[0][ok], b[0][ok]) and np.array_equal(a[1][ok], b[1][ok])
    finally:
        sim.use_point_machine(False)


def test_closed_form_eval_equals_hyper_dual_all_categories():
    """mix_eval (closed-form base + dipole terms, hyper-dual association on top) against the fully hyper-dual
    mix_eval_generic for every dipole / association category of tests/test_pcsaft_mix.py:17-36, with the segment
    numbers varied across the m = 2 clamp of the dipole term."""
    import cases

    rng = np.random.default_rng(1)
    reps = 100
    P = np.tile(np.array(cases.MIX_PARAMS, dtype=float).reshape(-1, 16), (reps, 1))
    K = np.tile(np.array(cases.MIX_KIJ), (reps, 1))
    n = P.shape[0]
    P[:, 0] *= rng.uniform(0.8, 1.3, n)
    P[:, 8] *= rng.uniform(0.6, 1.2, n)
    T = rng.uniform(200, 400, n)
    rho = np.exp(rng.uniform(np.log(1e-8), np.log(4e-3), (n, 2)))
    fast, generic = sim.mix_eval_pair(P, K, T, rho)
    err = np.abs(fast - generic) / np.maximum(np.abs(generic), 1e-300)
    assert err.max() < 1e-11 and np.median(err) < 1e-15


def test_closed_form_gradient_equals_dual_passes(golden_mix):
    """mix_point_grad_closed (closed-form coefficient adjoints + reverse sweep through mix_coef, the path of the
    non-associating states) against the forward-mode passes over all 19 input directions."""
    e = golden_mix["category_grid"]
    P, K, T, z = np.array(e["params"]).reshape(-1, 16), np.array(e["kij"]), np.array(e["T"]), np.array(e["z"])
    g = np.linspace(0.5, 1.5, len(T))
    try:
        for bubble in (True, False):
            rho, val, st, ev = sim.mix_point(P, K, T, z, bubble)
            ok = st == 0
            sim.mix_bwd_closed(False)
            a = sim.mix_point_bwd(P[ok], K[ok], T[ok], rho[ok], bubble, g[ok])
            sim.mix_bwd_closed(True)
            b = sim.mix_point_bwd(P[ok], K[ok], T[ok], rho[ok], bubble, g[ok])
            scale = np.abs(a).max(axis=1, keepdims=True)
            assert (np.abs(a - b) / scale).max() < 1e-10
            nonassoc = (P[ok][:, 6] + P[ok][:, 7] + P[ok][:, 14] + P[ok][:, 15]) == 0
            assert nonassoc.sum() >= 8 and (np.abs(a - b)[nonassoc] > 0).any()  # the closed-form path did run
    finally:
        sim.mix_bwd_closed(True)
