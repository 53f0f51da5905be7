"""TEST INFRASTRUCTURE ONLY -- CPU oracle solver, never imported by the product path.

Restates what the reference's native layer asks the un-vendored ``feos`` 0.6 crate to do
(`/root/reference/src/pcsaft.rs:82-129,150-214`, `src/gc_pcsaft.rs:103-171`): per state point,
find the coexisting densities (pure VLE, binary bubble/dew point) or the liquid density at
(T, p), and flag states for which no solution exists.  ``feos`` itself is absent from
/root/reference (Cargo.toml:18, no lock file), so this follows its published algorithm shape
(SURVEY.md Appendix B): ideal-gas start -> density iterations -> Maxwell / Newton updates, with
a spinodal-based restart, trivial-solution test and iteration caps defining the failed mask.
PARITY OF THE SOLVER DENSITIES THEMSELVES IS UNPINNED by the reference repo (it holds no golden
densities); what is pinned is the end-to-end README known-answer vector (README.md:26-29) and the
fact that every returned density is a root of the reference's own Helmholtz model to 1e-13.

All routines are vectorised over states with torch float64 on the CPU (all host threads) and
double as the `cpu_baseline` "port" in bench.py.
"""
import math

import torch

from . import eos

PI = math.pi
MAX_ETA = 0.5  # feos PC-SAFT max packing fraction used for rho_max (sigma based)


# ------------------------------------------------------------------ pure component helpers
def _pure_eval(params, T, rho):
    """a, da/drho, d2a/drho2 at fixed parameters (no graph kept)."""
    with torch.enable_grad():
        r = rho.detach().clone().requires_grad_(True)
        a = eos.pure_helmholtz(params.detach(), T.detach(), r)
        (a1,) = torch.autograd.grad(a.sum(), r, create_graph=True)
        (a2,) = torch.autograd.grad(a1.sum(), r)
    return a.detach(), a1.detach(), a2.detach()


def _pure_p_dp(params, T, rho):
    a, a1, a2 = _pure_eval(params, T, rho)
    return rho + rho * a1 - a, 1.0 + rho * a2, a, a1


def pure_max_density(params):
    m, sigma = params[:, 0], params[:, 1]
    return MAX_ETA / (PI / 6.0 * m * sigma**3)


def _spinodal_start(params, T, rho_max, n_grid=96, n_bisect=45):
    """Robust restart: bracket the mechanically unstable density range on a grid, bisect both
    spinodals, and return (ok, p_start, rho_sv, rho_sl).  ok=False <=> isotherm is monotonic
    (supercritical) on the grid."""
    N = T.shape[0]
    frac = (torch.arange(n_grid, dtype=torch.float64) + 0.5) / n_grid
    frac = frac**2 * 1.05  # denser at low density; reaches slightly above rho_max*1.0
    rho_g = rho_max[:, None] * frac[None, :]
    P = params[:, None, :].expand(N, n_grid, 8).reshape(-1, 8)
    Tg = T[:, None].expand(N, n_grid).reshape(-1)
    _, dp, _, _ = _pure_p_dp(P, Tg, rho_g.reshape(-1))
    neg = (dp.reshape(N, n_grid) < 0) | torch.isnan(dp.reshape(N, n_grid))
    ok = neg.any(dim=1) & ~neg[:, 0]
    idx = torch.arange(n_grid)[None, :].expand(N, n_grid)
    first = torch.where(neg, idx, n_grid).min(dim=1).values.clamp(1, n_grid - 1)
    last = torch.where(neg, idx, -1).max(dim=1).values.clamp(0, n_grid - 2)
    rows = torch.arange(N)

    def bisect(lo, hi, lo_is_stable):
        for _ in range(n_bisect):
            mid = 0.5 * (lo + hi)
            _, dpm, _, _ = _pure_p_dp(params, T, mid)
            stable = dpm > 0
            take_lo = stable if lo_is_stable else ~stable
            lo = torch.where(take_lo, mid, lo)
            hi = torch.where(take_lo, hi, mid)
        return lo, hi

    v_lo, _ = bisect(rho_g[rows, first - 1], rho_g[rows, first], True)  # stable side of vapour spinodal
    _, l_hi = bisect(rho_g[rows, last], rho_g[rows, last + 1], False)  # stable side of liquid spinodal
    p_sv, _, _, _ = _pure_p_dp(params, T, v_lo)
    p_sl, _, _, _ = _pure_p_dp(params, T, l_hi)
    p_start = 0.5 * (p_sl.clamp(min=0.0) + p_sv)
    ok = ok & (p_sv > 0) & (p_sv > p_sl)
    return ok, p_start, v_lo, l_hi


def _bracketed_density(params, T, p_target, lo, hi, n_iter=60):
    """Root of p(rho) = p_target on a stable branch [lo, hi] with p monotone: safeguarded Newton."""
    rho = 0.5 * (lo + hi)
    for _ in range(n_iter):
        p, dp, _, _ = _pure_p_dp(params, T, rho)
        below = p < p_target
        lo = torch.where(below, rho, lo)
        hi = torch.where(below, hi, rho)
        new = rho - (p - p_target) / dp
        bad = ~(new > lo) | ~(new < hi) | ~(dp > 0)
        rho = torch.where(bad, 0.5 * (lo + hi), new)
    return rho


def _maxwell_loop(params, T, rho_v, rho_l, max_iter=60, tol=1e-13):
    """feos-shaped pure VLE iteration: Maxwell pressure from the current
    densities, then one Newton step per phase; converged when both relative steps < tol."""
    N = T.shape[0]
    done = torch.zeros(N, dtype=torch.bool)
    bad = torch.zeros(N, dtype=torch.bool)
    for _ in range(max_iter):
        act = ~(done | bad)
        if not bool(act.any()):
            break
        P, Tt = params[act], T[act]
        rv, rl = rho_v[act], rho_l[act]
        p_v, dp_v, a_v, _ = _pure_p_dp(P, Tt, rv)
        p_l, dp_l, a_l, mu_l = _pure_p_dp(P, Tt, rl)
        dv = 1.0 / rv - 1.0 / rl
        p_new = -(a_v / rv - a_l / rl + torch.log(rv / rl)) / dv
        # negative Maxwell pressure: treat the vapour as ideal (always positive)
        p_id = rl * torch.exp(mu_l - p_l / rl)
        p_new = torch.where(p_new > 0, p_new, p_id)
        new_v = rv + (p_new - p_v) / dp_v
        new_l = rl + (p_new - p_l) / dp_l
        fail = (
            torch.isnan(new_v) | torch.isnan(new_l) | (new_v <= 0) | (new_l <= 0) | (dp_v <= 0) | (dp_l <= 0)
            | ((new_l / new_v - 1.0).abs() < 1e-5)
        )
        conv = ~fail & ((new_v - rv).abs() < tol * rv) & ((new_l - rl).abs() < tol * rl)
        ia = torch.nonzero(act).flatten()
        rho_v = rho_v.clone()
        rho_l = rho_l.clone()
        rho_v[ia] = torch.where(fail, rv, new_v)
        rho_l[ia] = torch.where(fail, rl, new_l)
        done[ia[conv]] = True
        bad[ia[fail]] = True
    return rho_v, rho_l, done


def pure_vle(params, temperature):
    """Coexisting (rho_V, rho_L) in 1/A^3 and failed mask; restates src/pcsaft.rs:82-103
    (PhaseEquilibrium::pure at fixed T with default options)."""
    params = params.detach().to(torch.float64)
    T = temperature.detach().to(torch.float64)
    N = T.shape[0]
    rho_max = pure_max_density(params)
    rho_v = torch.full((N,), float("nan"), dtype=torch.float64)
    rho_l = torch.full((N,), float("nan"), dtype=torch.float64)
    ok = torch.zeros(N, dtype=torch.bool)
    if N == 0:
        return rho_v, rho_l, ~ok

    # 1) ideal-gas start: p0 = rho kT exp(mu_res - p v) from a liquid-like state at 0.75 rho_max
    #    (feos: starting_pressure_ideal_gas_bubble), then density iterations at p0 with the liquid
    #    started at rho_max and the vapour at the ideal gas (feos: new_npt per phase)
    r0 = 0.75 * rho_max
    p0l, _, _, mu0 = _pure_p_dp(params, T, r0)
    p_start = r0 * torch.exp(mu0 - p0l / r0)
    good = torch.isfinite(p_start) & (p_start > 0)
    rv = torch.where(good, p_start, torch.ones_like(p_start) * 1e-10)
    rl = rho_max.clone()
    for _ in range(40):
        p_v, dp_v, _, _ = _pure_p_dp(params, T, rv)
        p_l, dp_l, _, _ = _pure_p_dp(params, T, rl)
        step_v = (p_start - p_v) / dp_v
        step_l = ((p_start - p_l) / dp_l).clamp(-0.075 * rho_max, 0.075 * rho_max)
        good = good & (dp_v > 0) & (dp_l > 0)
        rv = torch.where(good, (rv + step_v).clamp(min=0.05 * rv), rv)
        rl = torch.where(good, rl + step_l, rl)
    good = good & ((rl / rv - 1.0).abs() > 1e-5) & (rl > rv)
    if bool(good.any()):
        g = torch.nonzero(good).flatten()
        v, l, c = _maxwell_loop(params[g], T[g], rv[g], rl[g])
        rho_v[g[c]] = v[c]
        rho_l[g[c]] = l[c]
        ok[g[c]] = True

    # 2) spinodal-based restart for everything that did not converge
    rest = torch.nonzero(~ok).flatten()
    if rest.numel():
        P, Tt, rm = params[rest], T[rest], rho_max[rest]
        s_ok, p_s, r_sv, r_sl = _spinodal_start(P, Tt, rm)
        if bool(s_ok.any()):
            k = torch.nonzero(s_ok).flatten()
            Pk, Tk = P[k], Tt[k]
            v0 = _bracketed_density(Pk, Tk, p_s[k], torch.zeros_like(p_s[k]), r_sv[k])
            l0 = _bracketed_density(Pk, Tk, p_s[k], r_sl[k], 1.6 * rm[k])
            v, l, c = _maxwell_loop(Pk, Tk, v0, l0, max_iter=200)
            gi = rest[k]
            rho_v[gi[c]] = v[c]
            rho_l[gi[c]] = l[c]
            ok[gi[c]] = True
    return rho_v, rho_l, ~ok


def pure_liquid_density(params, temperature, pressure_pa):
    """Liquid-like density root of p(rho) = p at fixed T (src/pcsaft.rs:105-129:
    State::new_npt(.., DensityInitialization::Liquid)).  Newton from a high density with feos' step
    clamp (0.075 rho_max); a state that walks into the mechanically unstable range without finding
    a root, or does not converge in 50 iterations, is flagged failed."""
    params = params.detach().to(torch.float64)
    T = temperature.detach().to(torch.float64)
    p_t = pressure_pa.detach().to(torch.float64) / T / eos.PA_PER_RED
    rho_max = pure_max_density(params)
    rho = rho_max.clone()
    N = T.shape[0]
    done = torch.zeros(N, dtype=torch.bool)
    bad = torch.zeros(N, dtype=torch.bool)
    unstable_hits = torch.zeros(N, dtype=torch.int64)
    for _ in range(80):
        act = ~(done | bad)
        if not bool(act.any()):
            break
        p, dp, _, _ = _pure_p_dp(params, T, rho)
        step = (-(p - p_t) / dp).clamp(-0.075 * rho_max, 0.075 * rho_max)
        step = torch.maximum(step, -0.95 * rho)
        unstable = ~(dp > 0)
        unstable_hits = unstable_hits + (act & unstable).long()
        # unstable: move towards the branch that can hold the target pressure
        jump = torch.where(p > p_t, 0.5 * rho, torch.minimum(1.1 * rho, rho_max))
        new = torch.where(unstable, jump, rho + step)
        conv = act & ~unstable & ((new - rho).abs() <= 1e-13 * rho)  # the accepted step squares this
        rho = torch.where(act, new, rho)
        done = done | conv
        bad = bad | (act & (torch.isnan(new) | (unstable_hits > 20)))
    return rho, ~done


# ------------------------------------------------------------------ binary bubble / dew point
def _mix_eval(helm, T, rho):
    """a, grad a [N,2], Hessian a [N,2,2] in the partial densities."""
    with torch.enable_grad():
        r = rho.detach().clone().requires_grad_(True)
        a = helm(T, r)
        (g,) = torch.autograd.grad(a.sum(), r, create_graph=True)
        h0 = torch.autograd.grad(g[:, 0].sum(), r, retain_graph=True)[0]
        h1 = torch.autograd.grad(g[:, 1].sum(), r)[0]
    return a.detach(), g.detach(), torch.stack([h0.detach(), h1.detach()], dim=1)


def _mix_props(helm, T, rho):
    a, g, h = _mix_eval(helm, T, rho)
    rt = rho.sum(dim=1)
    p = rt - a + (rho * g).sum(dim=1)
    dp = 1.0 + (h * rho[:, :, None]).sum(dim=1)  # dp/drho_j = 1 + sum_i rho_i a_ij
    return a, g, h, p, dp


def _density_at_p(helm, T, x, p_t, rho0, rho_max, liquid, n_iter=40):
    """Total density of composition x at pressure p_t; Newton along the composition line."""
    rho = rho0.clone()
    for _ in range(n_iter):
        _, _, _, p, dpj = _mix_props(helm, T, rho[:, None] * x)
        dp = (dpj * x).sum(dim=1)
        step = -(p - p_t) / dp
        if liquid:
            step = step.clamp(-0.075 * rho_max, 0.075 * rho_max)
            new = torch.where(dp > 0, rho + step, torch.minimum(1.1 * rho, 1.2 * rho_max))
        else:
            new = torch.where(dp > 0, (rho + step).clamp(min=0.2 * rho), 0.5 * rho)
        new = torch.where(torch.isfinite(new), new, rho)
        done = bool(((new - rho).abs() <= 1e-14 * rho).all())
        rho = new
        if done:
            break
    return rho


def binary_bubble_dew(helm, max_density_fn, temperature, molefrac1, pressure_pa, bubble):
    """Binary bubble (given liquid x) or dew (given vapour y) point at fixed T.
    Restates PhaseEquilibrium::bubble_point/dew_point as called at src/pcsaft.rs:170-177,203-210
    and src/gc_pcsaft.rs:125-132,160-167.  Returns rho [N,4] = (rho_V1, rho_V2, rho_L1, rho_L2)
    in 1/A^3 (filter_binary order, src/pcsaft.rs:216-231) for ALL states plus the failed mask.

    Method: successive substitution on K-factors with a Newton update of ln p per sweep (the
    feos outer/inner structure), started from the caller's pressure, then polished by Newton on the
    three equilibrium conditions in (ln rho) variables to machine precision."""
    T = temperature.detach().to(torch.float64)
    z1 = molefrac1.detach().to(torch.float64)
    N = T.shape[0]
    z = torch.stack([z1, 1.0 - z1], dim=1)
    p = pressure_pa.detach().to(torch.float64) / T / eos.PA_PER_RED
    failed = torch.zeros(N, dtype=torch.bool)

    def rmax(x):
        return max_density_fn(x)

    w = z.clone()  # incipient-phase composition estimate
    # initial estimate of the incipient phase from a zero-pressure-like liquid (ideal vapour)
    if bubble:
        rl = _density_at_p(helm, T, z, p, rmax(z), rmax(z), True)
        _, g, _, p_l, _ = _mix_props(helm, T, rl[:, None] * z)
        part = z * rl[:, None] * torch.exp(g)  # ideal-vapour partial densities
        p = torch.where(torch.isfinite(part.sum(dim=1)) & (part.sum(dim=1) > 0), part.sum(dim=1), p)
        w = part / part.sum(dim=1, keepdim=True)
    else:
        # find x with y(x) = y for an ideal vapour over a liquid near zero pressure
        # i.e. the root in s = ln(x_1 / x_2) of  f(s) = s + mu_1(x) - mu_2(x) - ln(y_1 / y_2).  The fixed-point
        # form s <- s_t - (mu_1 - mu_2), even damped, oscillates without end for strongly associating pairs
        # (four-site components), so the root is bracketed and bisected (f grows with s for a stable liquid).
        s_t = torch.log(z[:, 0] / z[:, 1])

        def f_of(s):
            x1 = torch.sigmoid(s)
            x = torch.stack([x1, 1.0 - x1], dim=1)
            rl = _density_at_p(helm, T, x, torch.zeros_like(p), rmax(x), rmax(x), True, n_iter=12)
            _, g, _, _, _ = _mix_props(helm, T, rl[:, None] * x)
            return s + (g[:, 0] - g[:, 1]) - s_t

        lo = torch.full_like(s_t, -30.0)
        hi = torch.full_like(s_t, 30.0)
        f_lo = torch.full_like(s_t, float("nan"))  # residuals at the bracket ends, once an iterate has landed there
        f_hi = torch.full_like(s_t, float("nan"))
        side = torch.zeros_like(s_t)  # Illinois: which end was replaced last
        for it in range(60):
            mid = 0.5 * (lo + hi)
            if it >= 8:
                # regula falsi (Illinois variant) where both end residuals are known; bisection otherwise
                sec = (lo * f_hi - hi * f_lo) / (f_hi - f_lo)
                use = torch.isfinite(sec) & (sec > lo) & (sec < hi)
                mid = torch.where(use, sec, mid)
            f = f_of(mid)
            neg = f < 0
            # the retained end's residual is halved when the same end is replaced twice in a row
            f_hi = torch.where(neg & (side < 0), 0.5 * f_hi, f_hi)
            f_lo = torch.where(~neg & (side > 0), 0.5 * f_lo, f_lo)
            lo, f_lo = torch.where(neg, mid, lo), torch.where(neg, f, f_lo)
            hi, f_hi = torch.where(neg, hi, mid), torch.where(neg, f_hi, f)
            side = torch.where(neg, -torch.ones_like(side), torch.ones_like(side))
            if bool((((hi - lo) < 1e-9) | (f.abs() < 1e-11)).all()):
                break
        s = torch.where(f.abs() < 1e-11, mid, 0.5 * (lo + hi))
        x1 = torch.sigmoid(s)
        w = torch.stack([x1, 1.0 - x1], dim=1)
        rl = _density_at_p(helm, T, w, torch.zeros_like(p), rmax(w), rmax(w), True, n_iter=12)
        _, g, _, _, _ = _mix_props(helm, T, rl[:, None] * w)
        part = w * rl[:, None] * torch.exp(g)
        p = torch.where(torch.isfinite(part.sum(dim=1)) & (part.sum(dim=1) > 0), part.sum(dim=1), p)

    x_l = z if bubble else w
    y_v = w if bubble else z
    rho_l = _density_at_p(helm, T, x_l, p, rmax(x_l), rmax(x_l), True)
    rho_v = p.clone()

    def pack(rho_l, rho_v, x_l, y_v):
        if bubble:
            return torch.stack([torch.log(rho_l), torch.log(rho_v * y_v[:, 0]), torch.log(rho_v * y_v[:, 1])], dim=1)
        return torch.stack([torch.log(rho_v), torch.log(rho_l * x_l[:, 0]), torch.log(rho_l * x_l[:, 1])], dim=1)

    def sweeps(p, rho_l, rho_v, x_l, y_v):
        # successive substitution sweeps (moderate count; Newton polish follows)
        for _ in range(40):
            rho_l = _density_at_p(helm, T, x_l, p, rho_l, rmax(x_l), True, n_iter=6)
            rho_v = _density_at_p(helm, T, y_v, p, rho_v, rmax(y_v), False, n_iter=6)
            _, g_l, _, _, _ = _mix_props(helm, T, rho_l[:, None] * x_l)
            _, g_v, _, _, _ = _mix_props(helm, T, rho_v[:, None] * y_v)
            lnk = (g_l - torch.log(p / rho_l)[:, None]) - (g_v - torch.log(p / rho_v)[:, None])  # ln phi_L - ln phi_V
            # damped successive substitution (geometric mean of old and new): plain substitution oscillates
            # with growing amplitude for strongly non-ideal (cross-associating) pairs
            if bubble:
                new = x_l * torch.exp(lnk)
                s = new.sum(dim=1)
                upd = torch.sqrt(y_v * new / s[:, None])
                upd = upd / upd.sum(dim=1, keepdim=True)
                change = (upd - y_v).abs().max()
                y_v = upd
                p = p * torch.sqrt(s.clamp(0.5, 2.0))
            else:
                new = y_v * torch.exp(-lnk)
                s = new.sum(dim=1)
                upd = torch.sqrt(x_l * new / s[:, None])
                upd = upd / upd.sum(dim=1, keepdim=True)
                change = (upd - x_l).abs().max()
                x_l = upd
                p = p / torch.sqrt(s.clamp(0.5, 2.0))
            if bool(change < 1e-7) and bool(((s - 1.0).abs() < 1e-7).all()):
                break  # the Newton polish below takes it from here
        return pack(rho_l, rho_v, x_l, y_v)

    def newton(u):
        conv = torch.zeros(N, dtype=torch.bool)
        for _ in range(60):
            spec_rho = torch.exp(u[:, 0])[:, None] * z  # specified phase
            inc_rho = torch.exp(u[:, 1:3])
            a_s, g_s, h_s, p_s, dp_s = _mix_props(helm, T, spec_rho)
            a_i, g_i, h_i, p_i, dp_i = _mix_props(helm, T, inc_rho)
            F = torch.stack(
                [
                    g_s[:, 0] + torch.log(spec_rho[:, 0]) - g_i[:, 0] - torch.log(inc_rho[:, 0]),
                    g_s[:, 1] + torch.log(spec_rho[:, 1]) - g_i[:, 1] - torch.log(inc_rho[:, 1]),
                    p_s - p_i,
                ],
                dim=1,
            )
            J = torch.zeros((N, 3, 3), dtype=torch.float64)
            # d/d ln(rho_spec): rho_spec_j d/drho_j
            J[:, 0, 0] = (h_s[:, 0, :] * spec_rho).sum(dim=1) + 1.0
            J[:, 1, 0] = (h_s[:, 1, :] * spec_rho).sum(dim=1) + 1.0
            J[:, 2, 0] = (dp_s * spec_rho).sum(dim=1)
            for k in range(2):
                J[:, 0, 1 + k] = -(h_i[:, 0, k] * inc_rho[:, k]) - (1.0 if k == 0 else 0.0)
                J[:, 1, 1 + k] = -(h_i[:, 1, k] * inc_rho[:, k]) - (1.0 if k == 1 else 0.0)
                J[:, 2, 1 + k] = -dp_i[:, k] * inc_rho[:, k]
            good = torch.isfinite(J).all(dim=(1, 2)) & torch.isfinite(F).all(dim=1)
            Jg = torch.where(good[:, None, None], J, torch.eye(3, dtype=torch.float64).expand(N, 3, 3))
            Fg = torch.where(good[:, None], F, torch.zeros_like(F))
            try:
                du = torch.linalg.solve(Jg, -Fg[:, :, None])[:, :, 0]
            except RuntimeError:
                du = -torch.linalg.pinv(Jg) @ Fg[:, :, None]
                du = du[:, :, 0]
            scale = (1.0 / du.abs().max(dim=1).values.clamp(min=1e-300)).clamp(max=1.0)  # |du| <= 1
            du = du * scale[:, None]
            conv = good & (du.abs().max(dim=1).values < 1e-11)  # the accepted step squares this
            u = torch.where(conv[:, None] | ~good[:, None], u, u + du)
            if bool((conv | ~good).all()):
                break
        spec_rho = torch.exp(u[:, 0])[:, None] * z
        inc_rho = torch.exp(u[:, 1:3])
        if bubble:
            rho = torch.cat([inc_rho, spec_rho], dim=1)
        else:
            rho = torch.cat([spec_rho, inc_rho], dim=1)
        tot_v = rho[:, 0:2].sum(dim=1)
        tot_l = rho[:, 2:4].sum(dim=1)
        trivial = ((tot_l / tot_v - 1.0).abs() < 1e-5) & ((rho[:, 0] / tot_v - rho[:, 2] / tot_l).abs() < 1e-5)
        failed = ~conv | trivial | ~(tot_l > tot_v) | ~torch.isfinite(rho).all(dim=1)
        return rho, failed

    # stage 1: Newton straight from the ideal-vapour start (at low pressure it is already within the
    # convergence radius, and successive substitution can throw strongly non-ideal pairs off);
    # stage 2, for what stage 1 did not converge: successive substitution sweeps, then Newton
    rho, failed = newton(pack(rho_l, rho_v, x_l, y_v))
    if bool(failed.any()):
        rho2, failed2 = newton(sweeps(p, rho_l, rho_v, x_l, y_v))
        rho = torch.where(failed[:, None], rho2, rho)
        failed = failed & failed2
    return rho, failed


def mix_max_density(params):
    """x [N,2] -> rho_max for a PC-SAFT binary (eta_max / (pi/6 sum x_i m_i sigma_i^3))."""
    m, sigma = params[:, :, 0].detach(), params[:, :, 1].detach()

    def fn(x):
        return MAX_ETA / (PI / 6.0 * (x * m * sigma**3).sum(dim=1))

    return fn


def gc_max_density(tab):
    msig = (tab.m_seg.detach() * tab.sigma.detach() ** 3).sum(dim=2)  # [N,2]

    def fn(x):
        return MAX_ETA / (PI / 6.0 * (x * msig).sum(dim=1))

    return fn
