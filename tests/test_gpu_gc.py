"""GPU parity tests of the group-contribution bubble / dew point path, through the public API."""
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


def seg_args(golden_gc, requires_grad=False, device="cuda"):
    ident = list(golden_gc["segments"].keys())
    cols = tuple(torch.tensor([golden_gc["segments"][i][k] for i in ident], dtype=F64, device=device, requires_grad=requires_grad)
                 for k in range(8))
    return ident, cols


def test_state_functions(ft, golden_gc):
    e = golden_gc["test_grid"]
    n = len(e["T"])
    ident, par = seg_args(golden_gc)
    phi = torch.tensor([cases.GC_PHI] * n, dtype=F64, device="cuda")
    model = ft.GcPcSaftMix(ident, par, cases.GC_SEGMENTS, cases.GC_BONDS, cases.GC_KAB, phi)
    a, p, mu, v = model.derivatives(torch.tensor(e["T"], dtype=F64, device="cuda"), torch.tensor(e["rho"], dtype=F64, device="cuda"))
    # tests/test_gc_pcsaft.py:122-127
    assert np.abs(a.cpu().numpy() - e["a"]).max() < 1e-14
    assert np.abs(p.cpu().numpy() - e["p_red"]).max() < 1e-14
    assert np.abs(mu.cpu().numpy() - e["mu"]).max() < 1e-13
    assert (np.abs(v.cpu().numpy() - e["v"]) / np.abs(e["v"])).max() < 1e-11


@pytest.mark.parametrize("key,bubble", [("bubble", True), ("dew", False)])
def test_category_grid(ft, golden_gc, key, bubble):
    e = golden_gc["category_grid"]
    rows = e["category"]
    ge = e[key]
    n = len(rows)
    ident, par = seg_args(golden_gc, requires_grad=True)
    kab = [torch.tensor(k[2], dtype=F64, device="cuda", requires_grad=True) for k in cases.GC_KAB]
    recs = [(k[0], k[1], kv) for k, kv in zip(cases.GC_KAB, kab)]
    phi = torch.tensor([cases.GC_PHI] * n, dtype=F64, device="cuda", requires_grad=True)
    t = torch.tensor(e["T"], dtype=F64, device="cuda", requires_grad=True)
    z = torch.tensor(e["z"], dtype=F64, device="cuda")
    p0 = torch.tensor(e["p0"], dtype=F64, device="cuda")
    model = ft.GcPcSaftMix(ident, par, [cases.GC_SEGMENTS[i] for i in rows], [cases.GC_BONDS[i] for i in rows], recs, phi)
    val, nans = (model.bubble_point if bubble else model.dew_point)(t, z, p0)
    assert nans.cpu().tolist() == ge["failed"]
    assert (np.abs(val.detach().cpu().numpy() - ge["value"]) / np.abs(ge["value"])).max() < 1e-10
    val.sum().backward()
    for col in (0, 1, 3, 4, 5, 6, 7):
        want = np.array(ge["grad_seg%d" % col])
        assert np.abs(par[col].grad.cpu().numpy() - want).max() <= 1e-8 * np.abs(want).max() + 1e-300
    for j in range(2):
        assert abs(kab[j].grad.item() - ge["grad_kab%d" % j]) < 1e-8 * abs(ge["grad_kab%d" % j])
    assert (np.abs(t.grad.cpu().numpy() - ge["grad_T"]) / np.abs(ge["grad_T"])).max() < 1e-8
    assert torch.isfinite(phi.grad).all() and torch.isfinite(par[2].grad).all()
    # phi and epsilon_k gradients: the reference yields NaN (eps = 0 segment); check against finite differences
    with torch.no_grad():
        def total(phi_, eps_):
            pr = list(c.detach() for c in par)
            pr[2] = eps_
            m2 = ft.GcPcSaftMix(ident, tuple(pr), [cases.GC_SEGMENTS[i] for i in rows], [cases.GC_BONDS[i] for i in rows],
                                [(k[0], k[1], k[2]) for k in cases.GC_KAB], phi_)
            return (m2.bubble_point if bubble else m2.dew_point)(t.detach(), z, p0)[0]
        ph, ep = phi.detach(), par[2].detach()
        d = torch.zeros_like(ph)
        d[:, 1] = 1e-5
        fd = (total(ph + d, ep) - total(ph - d, ep)) / 2e-5
        # (h = 1e-5: truncation ~3e-8 relative, round-off floor ~1e-8 |p| for states that hardly depend on phi)
        assert ((fd - phi.grad[:, 1]).abs() <= 1e-6 * fd.abs() + 1e-7 * total(ph, ep).abs()).all()
        j = ident.index("CH2")
        de = torch.zeros_like(ep)
        de[j] = 1e-5 * ep[j]
        fd = (total(ph, ep + de).sum() - total(ph, ep - de).sum()) / (2 * de[j])
        assert abs(fd.item() - par[2].grad[j].item()) < 1e-6 * abs(fd.item())


@pytest.mark.parametrize("key,bubble", [("bubble", True), ("dew", False)])
def test_kab_only_gradient(ft, golden_gc, key, bubble):
    """A k_ab fit (only the binary segment records require a gradient) takes the k_ab-only backward kernel
    (gc_point_bwd_kab_kernel): same k_ab gradient as the reference (golden), per-state weights included."""
    e = golden_gc["category_grid"]
    rows = e["category"]
    ge = e[key]
    n = len(rows)
    ident, par = seg_args(golden_gc)
    kab = [torch.tensor(k[2], dtype=F64, device="cuda", requires_grad=True) for k in cases.GC_KAB]
    recs = [(k[0], k[1], kv) for k, kv in zip(cases.GC_KAB, kab)]
    phi = torch.tensor([cases.GC_PHI] * n, dtype=F64, device="cuda")
    t = torch.tensor(e["T"], dtype=F64, device="cuda")
    z = torch.tensor(e["z"], dtype=F64, device="cuda")
    p0 = torch.tensor(e["p0"], dtype=F64, device="cuda")
    segs, bonds = [cases.GC_SEGMENTS[i] for i in rows], [cases.GC_BONDS[i] for i in rows]
    model = ft.GcPcSaftMix(ident, par, segs, bonds, recs, phi)
    val, nans = (model.bubble_point if bubble else model.dew_point)(t, z, p0)
    val.sum().backward()  # broadcast upstream gradient
    for j in range(2):
        assert abs(kab[j].grad.item() - ge["grad_kab%d" % j]) < 1e-8 * abs(ge["grad_kab%d" % j])
    # per-state weights: against the general kernel (T requires a gradient there)
    w = torch.linspace(0.5, 1.5, n, dtype=F64, device="cuda")
    got = []
    for need_t in (False, True):
        for k in kab:
            k.grad = None
        tt = t.clone().requires_grad_(need_t)
        val, nans = (model.bubble_point if bubble else model.dew_point)(tt, z, p0)
        (val * w[~nans]).sum().backward()
        got.append([k.grad.item() for k in kab])
    for a, b in zip(*got):
        assert abs(a - b) <= 1e-12 * abs(b) and b != 0.0


def test_reference_configs(ft, golden_gc):
    """tests/test_gc_pcsaft.py:130-222."""
    cfg = cases.GC_POINT
    for name, bubble in (("test_bubble_point", True), ("test_dew_point", False)):
        g = golden_gc[name]
        ident, par = seg_args(golden_gc)
        kab = torch.tensor([cfg["kab"][0][2]], dtype=F64, device="cuda", requires_grad=True)
        recs = [(cfg["kab"][0][0], cfg["kab"][0][1], kab[0])]
        phi = torch.tensor([cfg["phi"]], dtype=F64, device="cuda")
        T = torch.tensor([cfg["T"]], dtype=F64, device="cuda", requires_grad=True)
        z = torch.tensor([cfg["z"]], dtype=F64, device="cuda", requires_grad=True)
        p0 = torch.tensor([cfg["p0"]], dtype=F64, device="cuda", requires_grad=True)
        model = ft.GcPcSaftMix(ident, par, [cfg["segments"]], [cfg["bonds"]], recs, phi)
        p, nans = (model.bubble_point if bubble else model.dew_point)(T, z, p0)
        assert not nans.any() and abs(p[0].item() - g["value"][0]) < 1e-8
        p[0].backward()
        assert abs(kab.grad[0].item() - g["grad_kab"][0]) < 1e-8 * abs(g["grad_kab"][0])
        assert abs(T.grad[0].item() - g["grad_T"][0]) < 1e-8 * abs(g["grad_T"][0])
        # raw native class
        rho, status = model.gc_pcsaft.bubble_point([cfg["T"]], [cfg["z"]], [cfg["p0"]]) if bubble else \
            model.gc_pcsaft.dew_point([cfg["T"]], [cfg["z"]], [cfg["p0"]])
        assert rho.shape == (1, 4) and not status.any() and rho[0, 2] + rho[0, 3] > rho[0, 0] + rho[0, 1]


def test_full_size_c5(ft, golden_gc):
    """BASELINE config C5: 1 M states over a small set of unique molecule pairs (non-associating rows
    0-4 of tests/test_gc_pcsaft.py:17-29), equilibrium conditions and shared k_ab gradient vs FD."""
    n = 1_000_000
    g = torch.Generator().manual_seed(9)
    ident, par = seg_args(golden_gc)
    which = torch.randint(0, 5, (n,), generator=g).tolist()
    pair_objs = [(cases.GC_SEGMENTS[i], cases.GC_BONDS[i]) for i in range(5)]
    segs = [pair_objs[w][0] for w in which]
    bonds = [pair_objs[w][1] for w in which]
    T = (150.0 + 100.0 * torch.rand(n, generator=g, dtype=F64)).cuda()
    z = (0.05 + 0.9 * torch.rand(n, generator=g, dtype=F64)).cuda()
    p0 = torch.full((n,), 1e5, dtype=F64, device="cuda")
    phi = torch.tensor([cases.GC_PHI], dtype=F64, device="cuda").expand(n, 2)
    kab = torch.tensor(0.03, dtype=F64, device="cuda", requires_grad=True)
    model = ft.GcPcSaftMix(ident, par, segs, bonds, [("CH3", "CH=O", kab), (">CH", "HCOO", -0.01)], phi)
    assert model._n_pairs == 5
    p, nans = model.bubble_point(T, z, p0)
    assert nans.float().mean().item() < 0.01 and torch.isfinite(p).all()
    p.sum().backward()
    grad = kab.grad.item()

    def total(k):
        m2 = ft.GcPcSaftMix(ident, par, segs, bonds, [("CH3", "CH=O", k), (">CH", "HCOO", -0.01)], phi)
        q, f = m2.bubble_point(T, z, p0)
        assert torch.equal(f, nans)
        return q.sum().item()

    fd = (total(0.03 + 1e-6) - total(0.03 - 1e-6)) / 2e-6
    assert abs(fd - grad) < 1e-5 * abs(fd)
    # equilibrium conditions through the raw densities of the native class on a slice
    k = 50_000
    rho, status = ft.GcPcSaft([(i, [c[j].item() for c in par]) for j, i in enumerate(ident)], segs[:k], bonds[:k],
                              [("CH3", "CH=O", 0.03), (">CH", "HCOO", -0.01)], phi[:k].cpu().numpy()).bubble_point(
        T[:k].cpu().numpy(), z[:k].cpu().numpy(), p0[:k].cpu().numpy())
    ok = torch.from_numpy(~status).cuda()
    m3 = ft.GcPcSaftMix(ident, par, segs[:k], bonds[:k], [("CH3", "CH=O", 0.03), (">CH", "HCOO", -0.01)], phi[:k])
    r = torch.zeros((k, 4), dtype=F64, device="cuda")
    r[ok] = torch.from_numpy(rho).cuda()
    _, pV, muV, _ = m3.derivatives(T[:k], r[:, 0:2].contiguous())
    _, pL, muL, _ = m3.derivatives(T[:k], r[:, 2:4].contiguous())
    assert ((pV - pL).abs() / r[:, 2:4].sum(dim=1))[ok].max().item() < 1e-12
    assert (muV + r[:, 0:2].log() - muL - r[:, 2:4].log())[ok].abs().max().item() < 1e-9


def test_state_order_is_a_permutation_and_changes_no_result(ft, golden_gc, monkeypatch):
    """csrc/order.cuh for the GC path: states are solved grouped by molecule pair, then (T bin, x).  Complete
    permutation; pressures, masks, densities and per-state gradients bit-identical to the input-order run; the
    shared k_ab gradient (atomic sums, order-dependent rounding) to 1e-12."""
    from feos_torch_b200 import _lib

    n = 100_000
    g = torch.Generator().manual_seed(12)
    ident, par = seg_args(golden_gc)
    which = torch.randint(0, 8, (n,), generator=g).tolist()  # rows 5-7 associate
    segs = [cases.GC_SEGMENTS[w] for w in which]
    bonds = [cases.GC_BONDS[w] for w in which]
    T = (150.0 + 100.0 * torch.rand(n, generator=g, dtype=F64)).cuda().requires_grad_(True)
    z = (0.05 + 0.9 * torch.rand(n, generator=g, dtype=F64)).cuda()
    p0 = torch.full((n,), 1e5, dtype=F64, device="cuda")
    phi = torch.tensor([cases.GC_PHI], dtype=F64, device="cuda").expand(n, 2).contiguous().requires_grad_(True)
    kab = torch.tensor(0.03, dtype=F64, device="cuda", requires_grad=True)
    model = ft.GcPcSaftMix(ident, par, segs, bonds, [("CH3", "CH=O", kab), (">CH", "HCOO", -0.01)], phi)

    order = torch.empty(n, dtype=torch.int32, device="cuda")
    for pair_only in (0, 1):
        _lib.call("feos_dev_gc_point_order", _lib.ptr(model._pair_index), model._n_pairs, _lib.ptr(T.detach()), _lib.ptr(z), n,
                  pair_only, _lib.ptr(order), torch.cuda.current_stream().cuda_stream)
        assert torch.equal(torch.sort(order.long()).values, torch.arange(n, device="cuda"))
        sorted_pairs = model._pair_index[order.long()]
        assert (sorted_pairs.diff() >= 0).all()
        if pair_only:  # stable: input order inside a pair
            assert (order.long().diff()[sorted_pairs.diff() == 0] > 0).all()

    def run(bubble):
        T.grad = phi.grad = kab.grad = None
        p, nans = (model.bubble_point if bubble else model.dew_point)(T, z, p0)
        p.sum().backward()
        return p.detach().clone(), nans.clone(), T.grad.clone(), phi.grad.clone(), kab.grad.clone()

    for bubble in (True, False):
        monkeypatch.setenv("FEOS_B200_STATE_ORDER", "1")
        a = run(bubble)
        monkeypatch.setenv("FEOS_B200_STATE_ORDER", "0")
        b = run(bubble)
        assert (~a[1]).float().mean().item() > 0.5
        for x, y in zip(a[:4], b[:4]):
            assert torch.equal(x, y)
        assert abs(a[4].item() - b[4].item()) <= 1e-12 * abs(b[4].item())
