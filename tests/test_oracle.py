"""Pins the oracle (CPU restatement, ``oracle/``) against
  * the README known-answer vector of the reference (README.md:26-29), and
  * ``tests/golden/*.json`` = outputs of the UNMODIFIED reference Python (imported through
    ``oracle/refshim`` in the build container, see ``oracle/make_golden.py``).
Tolerances follow BASELINE.json: rel 1e-10 on values, rel 1e-8 on gradients, identical mask.
"""
import os

import numpy as np
import pytest
import torch

import cases
from oracle import eos, models

F64 = torch.float64
# The oracle's binary solver is slow (Hessians by double backward in every iteration): by default the
# bubble/dew comparisons run on one state per category; FEOS_FULL_ORACLE=1 runs the whole grids
# (about 15 minutes; done when the golden files were generated).
FULL = bool(os.environ.get("FEOS_FULL_ORACLE"))


def close(a, b, rtol, atol=0.0):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    err = np.abs(a - b)
    tol = atol + rtol * np.abs(b)
    assert np.all(err <= tol), f"max violation {np.max(err - tol)} at {np.argmax(err - tol)}: {a.flat[np.argmax(err - tol)]} vs {b.flat[np.argmax(err - tol)]}"


def test_readme_known_answer():
    params = torch.tensor(cases.README_PARAMS, dtype=F64, requires_grad=True)
    model = models.PcSaftPure(params.repeat(5, 1))
    nans, vp = model.vapor_pressure(torch.tensor(cases.README_T, dtype=F64))
    assert not nans.any()
    for got, want in zip(vp.tolist(), cases.README_VP):
        assert f"{got:.4f}" == f"{want:.4f}"
    vp[0].backward()
    for got, want in zip(params.grad.tolist(), cases.README_GRAD):
        assert f"{got:.4e}" == f"{want:.4e}"


def test_pure_states_match_reference(golden_pure):
    for key in ("test_grid", "readme"):
        e = golden_pure[key]
        P = torch.tensor(e["params"], dtype=F64)
        T = torch.tensor(e["T"], dtype=F64)
        rho = torch.tensor(e["rho"], dtype=F64)
        a, p, dp = eos.pure_derivatives(P, T, rho)
        # tests/test_pcsaft_pure.py:59-61 uses abs 1e-10; the restatement is much closer than that
        close(a.detach(), e["a"], 1e-13, 1e-16)
        close(p.detach(), e["p_red"], 1e-13, 1e-16)
        close(dp.detach(), e["dp_red"], 1e-12, 1e-15)


def _check_pure_props(e, with_p):
    P0 = torch.tensor(e["params"], dtype=F64)
    T0 = torch.tensor(e["T"], dtype=F64)
    for prop in ("vapor_pressure", "liquid_density", "equilibrium_liquid_density"):
        g = e[prop]
        x = P0.clone().requires_grad_(True)
        t = T0.clone().requires_grad_(True)
        model = models.PcSaftPure(x)
        if prop == "liquid_density":
            pp = torch.tensor(e["p"], dtype=F64, requires_grad=True)
            nans, val = model.liquid_density(t, pp)
        else:
            nans, val = getattr(model, prop)(t)
        assert nans.tolist() == g["failed"]
        close(val.detach(), g["value"], 1e-10)
        val.sum().backward()
        gp = np.asarray(g["grad_params"])
        close(x.grad, gp, 1e-8, 1e-8 * np.abs(gp).max(axis=1, keepdims=True) * 1e-3)
        close(t.grad, g["grad_T"], 1e-8)
        if prop == "liquid_density":
            close(pp.grad, g["grad_p"], 1e-8)


def test_pure_properties_match_reference(golden_pure):
    for key in ("test_grid", "readme", "seeded_grid"):
        _check_pure_props(golden_pure[key], True)


def test_mix_states_match_reference(golden_mix):
    e = golden_mix["test_grid"]
    P = torch.tensor(e["params"], dtype=F64)
    K = torch.tensor(e["kij"], dtype=F64)
    T = torch.tensor(e["T"], dtype=F64)
    rho = torch.tensor(e["rho"], dtype=F64)
    a, p, mu, v = models.PcSaftMix(P, K).derivatives(T, rho)
    # tests/test_pcsaft_mix.py:119-124: abs 1e-14 on a, mu, p; 1e-11 on v.  The reference's inner
    # site-fraction iteration stops at 1e-10 (batch norm), so cross/induced rows agree to ~1e-11.
    close(a.detach(), e["a"], 0, 2e-12)
    close(p.detach(), e["p_red"], 0, 2e-12)
    close(mu.detach(), e["mu"], 0, 2e-9)
    close(v.detach(), e["v"], 1e-8)
    plain = [i for i, lab in enumerate(e["labels"]) if lab in ("np/np", "p/np", "np/p", "p/p", "a/np", "np/a")]
    close(a.detach()[plain], np.asarray(e["a"])[plain], 0, 1e-15)
    close(mu.detach()[plain], np.asarray(e["mu"])[plain], 0, 1e-13)
    close(v.detach()[plain], np.asarray(e["v"])[plain], 1e-12)


def _check_points(model_fn, e, leaves_fn, rtol_val=1e-10, rtol_grad=1e-7, per_state=(), skip_failed=False):
    """``leaves_fn(sel)`` builds the leaf tensors for the selected states; gradients of per-state leaves
    (names in ``per_state``) are compared row by row, shared leaves only on the full grid."""
    n = len(e["T"])
    for key, bubble in (("bubble", True), ("dew", False)):
        if key not in e:
            continue
        g = e[key]
        failed = np.asarray(g["failed"])
        if FULL or n <= 2:
            sel = np.arange(n)
        else:
            cat = np.asarray(e["category"])
            sel = np.array([np.nonzero(cat == c)[0][1] for c in sorted(set(cat.tolist()))])
        if skip_failed:
            sel = sel[~failed[sel]]  # states the (pre-fix) oracle solver could not converge
        whole = len(sel) == n
        T = torch.tensor(e["T"], dtype=F64)[sel]
        z = torch.tensor(e["z"], dtype=F64)[sel]
        p0 = torch.tensor(e["p0"], dtype=F64)[sel]
        leaves = leaves_fn(sel)
        t = T.clone().requires_grad_(True)
        model = model_fn(leaves, sel)
        val, nans = (model.bubble_point if bubble else model.dew_point)(t, z, p0)
        assert nans.tolist() == failed[sel].tolist(), key
        full = np.full(n, np.nan)
        full[~failed] = g["value"]
        close(val.detach(), full[sel][~failed[sel]], rtol_val)
        val.sum().backward()
        close(t.grad, np.asarray(g["grad_T"])[sel], rtol_grad, 1e-9)
        for name, leaf in leaves.items():
            want = g.get("grad_" + name)
            if want is None:
                continue
            want = np.asarray(want)
            if name in per_state:
                want = want[sel]
            elif not whole:
                continue
            if np.isnan(want).any():
                continue  # reference quirk: NaN gradient from sqrt(0) of absent eps=0 segments
            close(leaf.grad, want, rtol_grad, 1e-7 * np.abs(want).max() + 1e-12)


slow = pytest.mark.skipif(not FULL, reason="oracle binary solver on the category grids takes ~15 min; set FEOS_FULL_ORACLE=1")


def test_mix_reference_configs(golden_mix):
    """The reference's own bubble/dew test configurations (tests/test_pcsaft_mix.py:127-251)."""
    for name, key in (("test_bubble_point", "bubble"), ("test_dew_point", "dew")):
        g = golden_mix[name]
        ee = {k: g[k] for k in ("T", "z", "p0")}
        ee[key] = g

        def leaves1(sel):
            return {"params": torch.tensor(g["params"], dtype=F64, requires_grad=True),
                    "kij": torch.tensor(g["kij"], dtype=F64, requires_grad=True)}

        _check_points(lambda lv, sel: models.PcSaftMix(lv["params"], lv["kij"]), ee, leaves1, per_state=("params", "kij"))


@slow
def test_mix_points_match_reference(golden_mix):
    e = golden_mix["category_grid"]

    def leaves(sel):
        return {"params": torch.tensor(e["params"], dtype=F64)[sel].requires_grad_(True),
                "kij": torch.tensor(e["kij"], dtype=F64)[sel].requires_grad_(True)}

    # the golden masks were produced before the oracle's dew start was damped (it now converges a few
    # more states of the four-site categories): compare on the states the golden run converged
    _check_points(lambda lv, sel: models.PcSaftMix(lv["params"], lv["kij"]), e, leaves, per_state=("params", "kij"),
                  skip_failed=True)


def _gc_args(table, requires_grad=False):
    ident = list(table.keys())
    cols = tuple(torch.tensor([table[i][k] for i in ident], dtype=F64, requires_grad=requires_grad) for k in range(8))
    return ident, cols


def test_gc_states_match_reference(golden_gc):
    e = golden_gc["test_grid"]
    ident, par = _gc_args(golden_gc["segments"])
    n = len(cases.GC_SEGMENTS)
    phi = torch.tensor([cases.GC_PHI] * n, dtype=F64)
    model = models.GcPcSaftMix(ident, par, cases.GC_SEGMENTS, cases.GC_BONDS, cases.GC_KAB, phi)
    a, p, mu, v = model.derivatives(torch.tensor(e["T"], dtype=F64), torch.tensor(e["rho"], dtype=F64))
    close(a.detach(), e["a"], 0, 2e-12)
    close(p.detach(), e["p_red"], 0, 2e-12)
    close(mu.detach(), e["mu"], 0, 2e-9)
    close(v.detach(), e["v"], 1e-8)
    close(a.detach()[:5], np.asarray(e["a"])[:5], 0, 1e-15)


@slow
def test_gc_points_match_reference(golden_gc):
    e = golden_gc["category_grid"]
    rows = e["category"]
    segs = [cases.GC_SEGMENTS[i] for i in rows]
    bonds = [cases.GC_BONDS[i] for i in rows]

    def leaves(sel):
        ident, par = _gc_args(golden_gc["segments"], True)
        lv = {f"seg{j}": par[j] for j in range(8)}
        for j, k in enumerate(cases.GC_KAB):
            lv[f"kab{j}"] = torch.tensor(k[2], dtype=F64, requires_grad=True)
        lv["phi"] = torch.tensor([cases.GC_PHI] * len(sel), dtype=F64, requires_grad=True)
        return lv

    def model(lv, sel):
        ident = list(golden_gc["segments"].keys())
        par = tuple(lv[f"seg{j}"] for j in range(8))
        recs = [(k[0], k[1], lv[f"kab{j}"]) for j, k in enumerate(cases.GC_KAB)]
        return models.GcPcSaftMix(ident, par, [segs[i] for i in sel], [bonds[i] for i in sel], recs, lv["phi"])

    _check_points(model, e, leaves, per_state=("phi",))


# ---- the compiled per-state CPU solver (oracle/cpu_solver.cpp): bench.py's CPU baseline ---------------------
def test_compiled_solver_helmholtz_matches_reference(golden_pure):
    """Its Helmholtz energy (second-order duals in rho) against the reference Python's a, p, dp/drho."""
    from oracle import cpu_solver

    for key in ("test_grid", "readme"):
        e = golden_pure[key]
        rho = np.asarray(e["rho"])
        a, a1, a2 = cpu_solver.pure_helmholtz(np.asarray(e["params"]), np.asarray(e["T"]), rho)
        close(a, e["a"], 1e-13, 1e-16)
        close(rho - a + rho * a1, e["p_red"], 1e-13, 1e-16)
        close(1.0 + rho * a2, e["dp_red"], 1e-12, 1e-15)


def test_compiled_solver_matches_torch_solver(golden_pure):
    """Same algorithm as oracle/solver.py, one state at a time: identical masks, densities to 1e-12 -- on the golden
    grids (associating, dipolar and super-critical rows) and on a seeded C2-shaped grid."""
    from oracle import cpu_solver, solver

    sets = [(torch.tensor(golden_pure[k]["params"], dtype=F64), torch.tensor(golden_pure[k]["T"], dtype=F64))
            for k in ("test_grid", "readme", "seeded_grid")]
    g = torch.Generator().manual_seed(77)
    n = 400
    P = torch.zeros((n, 8), dtype=F64)
    P[:, 0] = 1.0 + 3.0 * torch.rand(n, generator=g, dtype=F64)
    P[:, 1] = 3.0 + torch.rand(n, generator=g, dtype=F64)
    P[:, 2] = 150.0 + 200.0 * torch.rand(n, generator=g, dtype=F64)
    tc = P[:, 2] * (0.6380 + 0.7070 * P[:, 0] - 0.06418 * P[:, 0] ** 2)
    sets.append((P, (0.5 + 0.7 * torch.rand(n, generator=g, dtype=F64)) * tc))  # tau up to 1.2: failed states too
    for P, T in sets:
        rv, rl, failed = cpu_solver.pure_vle(P, T)
        rv2, rl2, failed2 = solver.pure_vle(P, T)
        assert failed.tolist() == failed2.tolist()
        ok = ~failed
        close(rv[ok], rv2[ok], 1e-12)
        close(rl[ok], rl2[ok], 1e-12)
        pr = torch.full_like(T, 2e6)
        r, f = cpu_solver.pure_liquid_density(P, T, pr)
        r2, f2 = solver.pure_liquid_density(P, T, pr)
        assert f.tolist() == f2.tolist()
        close(r[~f], r2[~f], 1e-12)
    # README known answer through the compiled solver
    params = torch.tensor(cases.README_PARAMS, dtype=F64, requires_grad=True)
    nans, vp = models.PcSaftPure(params.repeat(5, 1), compiled=True).vapor_pressure(torch.tensor(cases.README_T, dtype=F64))
    for got, want in zip(vp.tolist(), cases.README_VP):
        assert f"{got:.4f}" == f"{want:.4f}"
