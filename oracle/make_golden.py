"""TEST INFRASTRUCTURE ONLY.  Generates ``tests/golden/*.json`` in the BUILD container.

Imports the UNMODIFIED reference Python from /root/reference through ``oracle.refshim`` (its native
``PcSaft`` / ``GcPcSaft`` classes replaced by the oracle CPU solver, ``oracle/native_stub.py``),
runs the reference's own test input sets (``tests/cases.py``) plus a few seeded grids through the
reference's public API and stores inputs + outputs (values, failed masks, autograd gradients) with
full float64 precision (Python ``repr`` round-trips exactly through JSON).

    python -m oracle.make_golden [pure|mix|gc ...]     # rewrites tests/golden/<name>.json

The group parameter table of the reference (tests/sauer2014_hetero.json) is re-emitted in a compact
form (identifier -> 8 numbers) because the GPU box has no /root/reference.
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import cases  # noqa: E402
from oracle import refshim  # noqa: E402

F64 = torch.float64
OUT = os.path.join(ROOT, "tests", "golden")


def tl(x):
    if isinstance(x, torch.Tensor):
        return x.detach().cpu().tolist()
    return x


def load_segments():
    with open(os.path.join(refshim.REFERENCE_ROOT, "tests", "sauer2014_hetero.json")) as f:
        recs = json.load(f)
    keys = ["m", "sigma", "epsilon_k", "mu", "kappa_ab", "epsilon_k_ab", "na", "nb"]
    return {r["identifier"]: [float(r["model_record"].get(k, 0.0)) for k in keys] for r in recs}


def seg_args(table, requires_grad=False):
    ident = list(table.keys())
    cols = [torch.tensor([table[i][k] for i in ident], dtype=F64, requires_grad=requires_grad) for k in range(8)]
    return ident, tuple(cols)


def pure_golden(ref):
    out = {}
    # --- reference test grid (tests/test_pcsaft_pure.py:9-88) and README (README.md:13-29)
    for name, rows, temps in (
        ("test_grid", cases.PURE_PARAMS, [300.0] * len(cases.PURE_PARAMS)),
        ("readme", [cases.README_PARAMS] * 5, cases.README_T),
    ):
        P = torch.tensor(rows, dtype=F64)
        T = torch.tensor(temps, dtype=F64)
        n = len(rows)
        rho = torch.full((n,), 0.001, dtype=F64)
        pr = torch.full((n,), 1e5, dtype=F64)
        entry = {"params": tl(P), "T": tl(T), "rho": tl(rho), "p": tl(pr)}
        a, p, dp = ref.PcSaftPure(P).derivatives(T, rho)
        entry["a"], entry["p_red"], entry["dp_red"] = tl(a), tl(p), tl(dp)
        entry["helmholtz"] = tl(ref.PcSaftPure(P).helmholtz_energy(T, rho))
        for prop in ("vapor_pressure", "liquid_density", "equilibrium_liquid_density"):
            x = P.clone().requires_grad_(True)
            t = T.clone().requires_grad_(True)
            pp = pr.clone().requires_grad_(True)
            model = ref.PcSaftPure(x)
            if prop == "liquid_density":
                nans, val = model.liquid_density(t, pp)
            else:
                nans, val = getattr(model, prop)(t)
            val.sum().backward()
            entry[prop] = {
                "failed": tl(nans), "value": tl(val), "grad_params": tl(x.grad), "grad_T": tl(t.grad),
                "grad_p": tl(pp.grad) if pp.grad is not None else None,
            }
        out[name] = entry

    # --- seeded grid: BASELINE config C2/C3 shaped, incl. a supercritical slice (failed mask)
    g = torch.Generator().manual_seed(1234)
    n = 96
    m = 1.0 + 3.0 * torch.rand(n, generator=g, dtype=F64)
    sigma = 3.0 + torch.rand(n, generator=g, dtype=F64)
    eps = 150.0 + 200.0 * torch.rand(n, generator=g, dtype=F64)
    tc = eps * (0.6380 + 0.7070 * m - 0.06418 * m * m)
    tau = 0.50 + 0.45 * torch.rand(n, generator=g, dtype=F64)
    tau[-12:] = 1.1 + 0.4 * torch.rand(12, generator=g, dtype=F64)  # supercritical -> failed
    P = torch.zeros((n, 8), dtype=F64)
    P[:, 0], P[:, 1], P[:, 2] = m, sigma, eps
    # a third of the sub-critical rows get 2B association, a few get a dipole
    P[0:28, 4] = torch.exp(torch.log(torch.tensor(0.003)) + torch.rand(28, generator=g, dtype=F64) * np.log(0.1 / 0.003))
    P[0:28, 5] = 1500.0 + 1500.0 * torch.rand(28, generator=g, dtype=F64)
    P[0:28, 6] = 1.0
    P[0:28, 7] = 1.0
    P[28:40, 3] = 1.0 + 1.5 * torch.rand(12, generator=g, dtype=F64)
    T = tau * tc
    T[0:28] = T[0:28] * 0.9  # association raises T_c; stay clearly sub-critical
    entry = {"params": tl(P), "T": tl(T)}
    x = P.clone().requires_grad_(True)
    t = T.clone().requires_grad_(True)
    nans, vp = ref.PcSaftPure(x).vapor_pressure(t)
    vp.sum().backward()
    entry["vapor_pressure"] = {"failed": tl(nans), "value": tl(vp), "grad_params": tl(x.grad), "grad_T": tl(t.grad)}
    x = P.clone().requires_grad_(True)
    t = T.clone().requires_grad_(True)
    nans2, rl = ref.PcSaftPure(x).equilibrium_liquid_density(t)
    rl.sum().backward()
    entry["equilibrium_liquid_density"] = {"failed": tl(nans2), "value": tl(rl), "grad_params": tl(x.grad), "grad_T": tl(t.grad)}
    # liquid density at p = max(1.05 p_sat, logU[1e5, 1e7]) for the converged rows, 1e6 Pa otherwise
    pr = torch.exp(np.log(1e5) + torch.rand(n, generator=g, dtype=F64) * np.log(100.0))
    psat = torch.full((n,), 0.0, dtype=F64)
    psat[~nans] = vp.detach()
    pr = torch.maximum(pr, 1.05 * psat)
    x = P.clone().requires_grad_(True)
    t = T.clone().requires_grad_(True)
    pp = pr.clone().requires_grad_(True)
    nans3, rho = ref.PcSaftPure(x).liquid_density(t, pp)
    rho.sum().backward()
    entry["p"] = tl(pr)
    entry["liquid_density"] = {"failed": tl(nans3), "value": tl(rho), "grad_params": tl(x.grad), "grad_T": tl(t.grad), "grad_p": tl(pp.grad)}
    out["seeded_grid"] = entry
    return out


def _point_entry(make_model, leaves, T, z, p0, bubble):
    """Run bubble/dew through the reference API and collect value, mask and gradients of ``leaves``."""
    for v in leaves.values():
        if v.grad is not None:
            v.grad = None
    t = T.clone().requires_grad_(True)
    model = make_model()
    fn = model.bubble_point if bubble else model.dew_point
    val, nans = fn(t, z.clone(), p0.clone())
    val.sum().backward()
    e = {"failed": tl(nans), "value": tl(val), "grad_T": tl(t.grad)}
    for k, v in leaves.items():
        e["grad_" + k] = tl(v.grad) if v.grad is not None else None
    return e


def mix_golden(ref, subset=None):
    """``subset``: indices into the category grid (reproducibility test: a slice of the grid through the same code);
    the random grid is skipped then."""
    out = {}
    P = torch.tensor(cases.MIX_PARAMS, dtype=F64)
    K = torch.tensor(cases.MIX_KIJ, dtype=F64)
    n = P.shape[0]
    T = torch.full((n,), 300.0, dtype=F64)
    rho = torch.tensor([[0.001, 0.002]] * n, dtype=F64)
    a, p, mu, v = ref.PcSaftMix(P, K).derivatives(T, rho)
    out["test_grid"] = {
        "labels": cases.MIX_LABELS, "params": tl(P), "kij": tl(K), "T": tl(T), "rho": tl(rho),
        "a": tl(a.flatten()), "p_red": tl(p.flatten()), "mu": tl(mu), "v": tl(v),
        "helmholtz_density": tl(ref.PcSaftMix(P, K).helmholtz_energy_density(T, rho).flatten()),
    }
    # bubble / dew for every category on a small (T, z) grid, gradients w.r.t. params and kij
    Tg = [180.0, 230.0, 280.0]
    zg = [0.2, 0.5, 0.8]
    rows, temps, zs = [], [], []
    for i in range(n):
        for t in Tg:
            for z in zg:
                rows.append(i)
                temps.append(t)
                zs.append(z)
    if subset is not None:
        rows, temps, zs = ([v[i] for i in subset] for v in (rows, temps, zs))
    idx = torch.tensor(rows)
    Pg, Kg = P[idx], K[idx]
    Tt = torch.tensor(temps, dtype=F64)
    zt = torch.tensor(zs, dtype=F64)
    p0 = torch.full_like(Tt, 1e5)
    entry = {"params": tl(Pg), "kij": tl(Kg), "T": tl(Tt), "z": tl(zt), "p0": tl(p0), "category": rows}
    for bubble in (True, False):
        x = Pg.clone().requires_grad_(True)
        k = Kg.clone().requires_grad_(True)
        entry["bubble" if bubble else "dew"] = _point_entry(lambda: ref.PcSaftMix(x, k), {"params": x, "kij": k}, Tt, zt, p0, bubble)
    out["category_grid"] = entry

    # seeded random (category, T, z) states: masks, values and gradients away from the test grid's nodes.  The
    # liquid / vapour density ratio of the oracle's solution is stored beside the mask: converged-or-failed is
    # implementation-defined next to a critical point, the tests exempt ratio < 1.5.
    g = torch.Generator().manual_seed(5)
    m = 10 * n if subset is None else 0
    cat = torch.arange(m) % n
    Tr = 150.0 + 140.0 * torch.rand(m, generator=g, dtype=F64)
    zr = 0.02 + 0.96 * torch.rand(m, generator=g, dtype=F64)
    p0 = torch.full_like(Tr, 1e5)
    Pr, Kr = P[cat], K[cat]
    entry = {"params": tl(Pr), "kij": tl(Kr), "T": tl(Tr), "z": tl(zr), "p0": tl(p0), "category": tl(cat)}
    from oracle import eos as oeos, solver as osolver
    for bubble in (True, False) if m else ():
        x = Pr.clone().requires_grad_(True)
        k = Kr.clone().requires_grad_(True)
        e = _point_entry(lambda: ref.PcSaftMix(x, k), {"params": x, "kij": k}, Tr, zr, p0, bubble)
        rho, failed = osolver.binary_bubble_dew(lambda T_, r_: oeos.mix_helmholtz(Pr, Kr, T_, r_), osolver.mix_max_density(Pr),
                                                Tr, zr, p0, bubble)
        assert failed.tolist() == e["failed"]
        e["rho_ratio"] = tl(rho[:, 2:4].sum(dim=1) / rho[:, 0:2].sum(dim=1))
        entry["bubble" if bubble else "dew"] = e
    if m:
        out["random_grid"] = entry

    # the reference's own bubble / dew test configurations (tests/test_pcsaft_mix.py:127-251)
    for name, cfg, bubble in (("test_bubble_point", cases.MIX_BUBBLE, True), ("test_dew_point", cases.MIX_DEW, False)):
        x = torch.tensor([cfg["params"]], dtype=F64, requires_grad=True)
        k = torch.tensor([cfg["kij"]], dtype=F64, requires_grad=True)
        Tt = torch.tensor([cfg["T"]], dtype=F64)
        zt = torch.tensor([cfg["z"]], dtype=F64)
        p0 = torch.tensor([cfg["p0"]], dtype=F64)
        e = _point_entry(lambda: ref.PcSaftMix(x, k), {"params": x, "kij": k}, Tt, zt, p0, bubble)
        e.update({"params": tl(x), "kij": tl(k), "T": tl(Tt), "z": tl(zt), "p0": tl(p0)})
        out[name] = e
    return out


def gc_golden(ref, table, subset=None):
    out = {"segments": table}
    n = len(cases.GC_SEGMENTS)
    T = torch.full((n,), 300.0, dtype=F64)
    rho = torch.tensor([[0.001, 0.002]] * n, dtype=F64)
    phi = torch.tensor([cases.GC_PHI] * n, dtype=F64)
    ident, par = seg_args(table)
    model = ref.GcPcSaftMix(ident, par, cases.GC_SEGMENTS, cases.GC_BONDS, cases.GC_KAB, phi)
    a, p, mu, v = model.derivatives(T, rho)
    out["test_grid"] = {"T": tl(T), "rho": tl(rho), "a": tl(a.flatten()), "p_red": tl(p.flatten()), "mu": tl(mu), "v": tl(v)}

    # bubble / dew for the non-associating and associating categories, small (T, z) grid
    Tg = [200.0, 260.0]
    zg = [0.3, 0.7]
    rows, temps, zs = [], [], []
    for i in range(n):
        for t in Tg:
            for z in zg:
                rows.append(i)
                temps.append(t)
                zs.append(z)
    if subset is not None:
        rows, temps, zs = ([v[i] for i in subset] for v in (rows, temps, zs))
    segs = [cases.GC_SEGMENTS[i] for i in rows]
    bonds = [cases.GC_BONDS[i] for i in rows]
    Tt = torch.tensor(temps, dtype=F64)
    zt = torch.tensor(zs, dtype=F64)
    p0 = torch.full_like(Tt, 1e5)
    entry = {"category": rows, "T": tl(Tt), "z": tl(zt), "p0": tl(p0), "phi": cases.GC_PHI, "kab": [list(k) for k in cases.GC_KAB]}
    for bubble in (True, False):
        ident, par = seg_args(table, requires_grad=True)
        kab = [torch.tensor(k[2], dtype=F64, requires_grad=True) for k in cases.GC_KAB]
        recs = [(k[0], k[1], kv) for k, kv in zip(cases.GC_KAB, kab)]
        ph = torch.tensor([cases.GC_PHI] * len(rows), dtype=F64, requires_grad=True)
        leaves = {f"seg{j}": par[j] for j in range(8)}
        leaves.update({f"kab{j}": kab[j] for j in range(len(kab))})
        leaves["phi"] = ph
        entry["bubble" if bubble else "dew"] = _point_entry(
            lambda: ref.GcPcSaftMix(ident, par, segs, bonds, recs, ph), leaves, Tt, zt, p0, bubble
        )
    out["category_grid"] = entry

    # the reference's own GC bubble / dew configuration (tests/test_gc_pcsaft.py:130-222)
    cfg = cases.GC_POINT
    for name, bubble in (("test_bubble_point", True), ("test_dew_point", False)):
        ident, par = seg_args(table)
        kab = torch.tensor([cfg["kab"][0][2]], dtype=F64, requires_grad=True)
        recs = [(cfg["kab"][0][0], cfg["kab"][0][1], kab[0])]
        ph = torch.tensor([cfg["phi"]], dtype=F64, requires_grad=True)
        Tt = torch.tensor([cfg["T"]], dtype=F64)
        zt = torch.tensor([cfg["z"]], dtype=F64)
        p0 = torch.tensor([cfg["p0"]], dtype=F64)
        e = _point_entry(
            lambda: ref.GcPcSaftMix(ident, par, [cfg["segments"]], [cfg["bonds"]], recs, ph),
            {"kab": kab, "phi": ph}, Tt, zt, p0, bubble,
        )
        out[name] = e
    return out


def state_grad_golden(ref, table):
    """Autograd gradients of the reference's public state functions `derivatives` (pcsaft_pure.py:180-182,
    pcsaft_mix.py:395-420, gc_pcsaft.py:443-468) with respect to parameters, k_ij / k_ab, phi, T and the
    densities, for a weighted sum of the outputs (weights stored)."""
    out = {}
    rng = np.random.default_rng(7)
    # pure
    P = torch.tensor(cases.PURE_PARAMS, dtype=F64)
    n = P.shape[0]
    T = torch.linspace(280.0, 330.0, n, dtype=F64)
    rho = torch.linspace(0.0008, 0.012, n, dtype=F64)
    w = torch.tensor(rng.uniform(0.5, 1.5, (n, 3)))
    x, t, r = P.clone().requires_grad_(True), T.clone().requires_grad_(True), rho.clone().requires_grad_(True)
    a, p, dp = ref.PcSaftPure(x).derivatives(t, r)
    (w[:, 0] * a + w[:, 1] * p + w[:, 2] * dp).sum().backward()
    out["pure"] = {"params": tl(P), "T": tl(T), "rho": tl(rho), "w": tl(w), "a": tl(a), "p": tl(p), "dp": tl(dp),
                   "grad_params": tl(x.grad), "grad_T": tl(t.grad), "grad_rho": tl(r.grad)}
    # binary mixtures (all association / dipole categories of tests/test_pcsaft_mix.py:17-36)
    P = torch.tensor(cases.MIX_PARAMS, dtype=F64)
    K = torch.tensor(cases.MIX_KIJ, dtype=F64)
    n = P.shape[0]
    T = torch.linspace(290.0, 320.0, n, dtype=F64)
    rho = torch.stack([torch.linspace(0.0008, 0.004, n, dtype=F64), torch.linspace(0.003, 0.001, n, dtype=F64)], dim=1)
    w = torch.tensor(rng.uniform(0.5, 1.5, (n, 6)))
    x, k, t, r = (v.clone().requires_grad_(True) for v in (P, K, T, rho))
    a, p, mu, v = ref.PcSaftMix(x, k).derivatives(t, r)
    (w[:, 0] * a + w[:, 1] * p + (w[:, 2:4] * mu).sum(1) + (w[:, 4:6] * v).sum(1)).sum().backward()
    out["mix"] = {"T": tl(T), "rho": tl(rho), "w": tl(w), "grad_params": tl(x.grad), "grad_kij": tl(k.grad),
                  "grad_T": tl(t.grad), "grad_rho": tl(r.grad)}
    # group contribution (tests/test_gc_pcsaft.py:17-44)
    ident, par = seg_args(table, requires_grad=True)
    n = len(cases.GC_SEGMENTS)
    kab = [torch.tensor(kk[2], dtype=F64, requires_grad=True) for kk in cases.GC_KAB]
    phi = torch.tensor([cases.GC_PHI] * n, dtype=F64, requires_grad=True)
    T = torch.linspace(290.0, 320.0, n, dtype=F64).requires_grad_(True)
    rho = torch.stack([torch.linspace(0.0008, 0.002, n, dtype=F64), torch.linspace(0.002, 0.001, n, dtype=F64)], dim=1).requires_grad_(True)
    w = torch.tensor(rng.uniform(0.5, 1.5, (n, 6)))
    model = ref.GcPcSaftMix(ident, par, cases.GC_SEGMENTS, cases.GC_BONDS, [(a_, b_, kv) for (a_, b_, _), kv in zip(cases.GC_KAB, kab)], phi)
    a, p, mu, v = model.derivatives(T, rho)
    (w[:, 0] * a + w[:, 1] * p + (w[:, 2:4] * mu).sum(1) + (w[:, 4:6] * v).sum(1)).sum().backward()
    out["gc"] = {"T": tl(T), "rho": tl(rho), "w": tl(w), "grad_seg": [tl(c.grad) for c in par], "grad_kab": [kk.grad.item() for kk in kab],
                 "grad_phi": tl(phi.grad), "grad_T": tl(T.grad), "grad_rho": tl(rho.grad)}
    return out


def main():
    if not refshim.reference_available():
        raise SystemExit("reference sources not found; golden vectors can only be made in the build container")
    ref = refshim.load_reference()
    os.makedirs(OUT, exist_ok=True)
    table = load_segments()
    which = sys.argv[1:] or ["pure", "mix", "gc"]
    makers = {"pure": lambda: pure_golden(ref), "mix": lambda: mix_golden(ref), "gc": lambda: gc_golden(ref, table),
              "state_grad": lambda: state_grad_golden(ref, table)}
    for name in which:
        data = makers[name]()
        with open(os.path.join(OUT, f"{name}.json"), "w") as f:
            json.dump(data, f, indent=0)
        print("wrote", name)


if __name__ == "__main__":
    main()
