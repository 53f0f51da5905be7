"""TEST INFRASTRUCTURE ONLY -- CPU oracle, never imported by the product path.

Restatement (torch float64, CPU, vectorised over states) of the PC-SAFT residual Helmholtz
models of the reference and of the closed-form Newton-step property formulas that carry its
gradients.  Every function cites the reference lines it follows.  Derivatives are obtained with
``torch.autograd`` (reverse mode, ``create_graph=True``) instead of the reference's dual-number
classes (`feos_torch/dual.py`, `feos_torch/dual_torch.py`) -- a deliberately different mechanism
for the same mathematical quantities, so that the oracle is an independent check.

Pinned by: README.md:26-29 known-answer vectors (tests/test_oracle.py) and, in the build
container, direct comparison with the unmodified reference Python imported through
``oracle/refshim`` (tests/test_oracle_vs_reference.py; fixtures in tests/golden/).

Units (reference `pcsaft_pure.py:196-199,215`): T [K], densities [1/Angstrom^3], a = A/(V kT)
[1/Angstrom^3], p = P/(kT) [1/Angstrom^3].
"""
import math

import torch

KB = 1.380649e-23
NAV = 6.02214076e23
ANGSTROM = 1.0e-10
# mu[D]^2 -> K Angstrom^3:  1e-19 J/K / kB   (pcsaft_pure.py:95-100)
DIPOLE_SCALE = 1e-19 / KB
PA_PER_RED = KB / ANGSTROM**3  # p[Pa] = p_red * T * PA_PER_RED       (pcsaft_pure.py:215)
KMOL_M3_PER_RED = 1.0 / (1.0e3 * NAV * ANGSTROM**3)  # rho[kmol/m3] = rho_red * this  (:199)

# universal model constants (pcsaft_pure.py:10-86)
_A = torch.tensor(
    [
        [0.91056314451539, -0.30840169182720, -0.09061483509767],
        [0.63612814494991, 0.18605311591713, 0.45278428063920],
        [2.68613478913903, -2.50300472586548, 0.59627007280101],
        [-26.5473624914884, 21.4197936296668, -1.72418291311787],
        [97.7592087835073, -65.2558853303492, -4.13021125311661],
        [-159.591540865600, 83.3186804808856, 13.7766318697211],
        [91.2977740839123, -33.7469229297323, -8.67284703679646],
    ],
    dtype=torch.float64,
)
_B = torch.tensor(
    [
        [0.72409469413165, -0.57554980753450, 0.09768831158356],
        [2.23827918609380, 0.69950955214436, -0.25575749816100],
        [-4.00258494846342, 3.89256733895307, -9.15585615297321],
        [-21.00357681484648, -17.21547164777212, 20.64207597439724],
        [26.8556413626615, 192.6722644652495, -38.80443005206285],
        [206.5513384066188, -161.8264616487648, 93.6267740770146],
        [-355.60235612207947, -165.2076934555607, -29.66690558514725],
    ],
    dtype=torch.float64,
)
_AD = torch.tensor(
    [
        [0.30435038064, 0.95346405973, -1.16100802773],
        [-0.13585877707, -1.83963831920, 4.52586067320],
        [1.44933285154, 2.01311801180, 0.97512223853],
        [0.35569769252, -7.37249576667, -12.2810377713],
        [-2.06533084541, 8.23741345333, 5.93975747420],
    ],
    dtype=torch.float64,
)
_BD = torch.tensor(
    [
        [0.21879385627, -0.58731641193, 3.48695755800],
        [-1.18964307357, 1.24891317047, -14.9159739347],
        [1.16268885692, -0.50852797392, 15.3720218600],
        [0.0, 0.0, 0.0],
        [0.0, 0.0, 0.0],
    ],
    dtype=torch.float64,
)
_CD = torch.tensor(
    [
        [-0.06467735252, -0.95208758351, -0.62609792333],
        [0.19758818347, 2.99242575222, 1.29246858189],
        [-0.80875619458, -2.38026356489, 1.65427830900],
        [0.69028490492, -0.27012609786, -3.43967436378],
    ],
    dtype=torch.float64,
)

PI = math.pi


def _powers(x, n):
    """[x^0, ..., x^(n-1)] stacked on a new last axis."""
    out = [torch.ones_like(x)]
    for _ in range(n - 1):
        out.append(out[-1] * x)
    return torch.stack(out, dim=-1)


def _series(coef, w1, w2, etas):
    """sum_i (c_i0 + w1 c_i1 + w2 c_i2) eta^i  -- shared shape of all PC-SAFT integrals."""
    n = coef.shape[0]
    c = coef[:, 0] + w1[..., None] * coef[:, 1] + w2[..., None] * coef[:, 2]
    return (c * etas[..., :n]).sum(dim=-1)


def _c1(mbar, eta):
    """Compressibility factor term of the dispersion contribution (pcsaft_pure.py:134-140)."""
    e2 = eta * eta
    om = 1.0 - eta
    return 1.0 / (
        1.0
        + mbar * (8.0 * eta - 2.0 * e2) / om**4
        + (1.0 - mbar) * (20.0 * eta - 27.0 * e2 + 12.0 * e2 * eta - 2.0 * e2 * e2) / (om * om * (2.0 - eta) ** 2)
    )


def _site_term(x):
    return x.log() - 0.5 * x + 0.5


# --------------------------------------------------------------------------------------------
# pure component                                                        pcsaft_pure.py:89-182
# --------------------------------------------------------------------------------------------
def pure_helmholtz(params, temperature, density):
    """a(T, rho) of a pure PC-SAFT fluid.  Follows pcsaft_pure.py:90-103 (parameter unpacking,
    reduced squared dipole moment) and :106-178 (HS, HC, dispersion, dipole, 2-site association).
    params [N,8] = m, sigma, epsilon_k, mu, kappa_ab, epsilon_k_ab, na, nb."""
    m, sigma, eps, mu, kappa, eps_ab, na, nb = params.unbind(dim=-1)
    T, rho = temperature, density
    s3 = sigma**3
    e = eps / T
    d = sigma * (1.0 - 0.12 * torch.exp(-3.0 * e))
    eta = PI / 6.0 * m * rho * d**3
    om = 1.0 - eta
    etas = _powers(eta, 7)

    hs = m * rho * (4.0 * eta - 3.0 * eta * eta) / (om * om)
    hc = -rho * (m - 1.0) * torch.log((1.0 - 0.5 * eta) / om**3)

    m1 = (m - 1.0) / m
    m2 = (m - 2.0) / m
    i1 = _series(_A, m1, m1 * m2, etas)
    i2 = _series(_B, m1, m1 * m2, etas)
    disp = -PI * rho * rho * m * m * e * s3 * (2.0 * i1 + _c1(m, eta) * i2 * m * e)

    # dipole (Gross & Vrabec), mu^2 factored out so that mu = 0 is regular (pcsaft_pure.py:145-160)
    mu2 = mu * mu / (m * s3 * eps) * DIPOLE_SCALE * e * s3
    mc = m.clamp(max=2.0)
    w1 = (mc - 1.0) / mc
    w2 = w1 * (mc - 2.0) / mc
    j1 = _series(_AD, w1, w2, etas) + e * _series(_BD, w1, w2, etas)
    j2 = _series(_CD, w1, w2, etas)
    phi2 = -rho * rho * j1 / s3 * PI
    phi3 = -rho * rho * rho * j2 / s3 * (4.0 / 3.0 * PI * PI)
    dipole = phi2 * phi2 * mu2 * mu2 / (phi2 - phi3 * mu2)

    # association, two site types A/B with multiplicities na/nb (pcsaft_pure.py:163-176)
    k = eta / om
    delta = (1.0 + k * (1.5 + 0.5 * k)) / om * (torch.exp(eps_ab / T) - 1.0) * s3 * kappa
    ra, rb = na * rho, nb * rho
    aux = 1.0 + (ra - rb) * delta
    root = torch.sqrt(aux * aux + 4.0 * rb * delta)
    xa = 2.0 / (root + 1.0 + (rb - ra) * delta)
    xb = 2.0 / (root + 1.0 - (rb - ra) * delta)
    assoc = ra * _site_term(xa) + rb * _site_term(xb)
    return hs + hc + disp + dipole + assoc


def _grad(y, x, create_graph=True):
    (g,) = torch.autograd.grad(y.sum(), x, create_graph=create_graph)
    return g


def pure_derivatives(params, temperature, density):
    """(a, p/kT, d(p/kT)/drho) -- pcsaft_pure.py:180-182, via autograd instead of Dual3."""
    rho = density.detach().clone().requires_grad_(True)
    a = pure_helmholtz(params, temperature, rho)
    a1 = _grad(a, rho)
    a2 = _grad(a1, rho)
    r = density
    return _swap(a, rho, r), r - _swap(a, rho, r) + r * _swap(a1, rho, r), 1.0 + r * _swap(a2, rho, r)


def _swap(y, leaf, value):
    # derivatives were taken with respect to a detached leaf; values are identical, and the
    # reference treats the solver densities as constants anyway (torch.from_numpy, :189,205).
    return y


def pure_vapor_pressure_from_densities(params, temperature, rho_v, rho_l):
    """Maxwell/Newton-step saturation pressure in Pa at fixed densities (pcsaft_pure.py:210-215)."""
    a_l = pure_helmholtz(params, temperature, rho_l) / rho_l
    a_v = pure_helmholtz(params, temperature, rho_v) / rho_v
    p = -(a_v - a_l + torch.log(rho_v / rho_l)) / (1.0 / rho_v - 1.0 / rho_l)
    return p * temperature * PA_PER_RED


def pure_liquid_density_from_density(params, temperature, pressure, rho):
    """Newton-step liquid density in kmol/m^3 at a fixed density (pcsaft_pure.py:196-199)."""
    p_target = pressure / temperature / PA_PER_RED
    _, p, dp = pure_derivatives(params, temperature, rho)
    return (rho - (p - p_target) / dp) * KMOL_M3_PER_RED


def pure_equilibrium_liquid_density_from_densities(params, temperature, rho_v, rho_l):
    """Newton-step saturated liquid density in kmol/m^3 (pcsaft_pure.py:226-233)."""
    a_l, p_l, dp_l = pure_derivatives(params, temperature, rho_l)
    a_l = a_l / rho_l
    a_v = pure_helmholtz(params, temperature, rho_v) / rho_v
    p = -(a_v - a_l + torch.log(rho_v / rho_l)) / (1.0 / rho_v - 1.0 / rho_l)
    return (rho_l - (p_l - p) / dp_l) * KMOL_M3_PER_RED


# --------------------------------------------------------------------------------------------
# shared mixture pieces
# --------------------------------------------------------------------------------------------
def _hard_sphere(z0, z1, z2, z3):
    """BMCSL hard-sphere term (pcsaft_mix.py:56-60, gc_pcsaft.py:149-153)."""
    om = 1.0 - z3
    z23 = z2 / z3
    return (6.0 / PI) * (3.0 * z1 * z2 / om + z2 * z2 * z23 / (om * om) + (z2 * z23 * z23 - z0) * torch.log(om))


def _dispersion_tail(rho1mix, rho2mix, mbar, z3):
    """pcsaft_mix.py:92-106 / gc_pcsaft.py:196-210."""
    etas = _powers(z3, 7)
    w1 = (mbar - 1.0) / mbar
    w2 = w1 * (mbar - 2.0) / mbar
    i1 = _series(_A, w1, w2, etas)
    i2 = _series(_B, w1, w2, etas)
    return (-2.0 * rho1mix * i1 - rho2mix * _c1(mbar, z3) * i2 * mbar) * PI


def _dipole_binary(m, sigma, eps, mu2_term, T, rho, z3):
    """Gross-Vrabec dipole term for n components given the per-component factor ``mu2_term``
    (pcsaft_mix.py:156-208, gc_pcsaft.py:255-307).  Returns 0 where no component is dipolar."""
    n = m.shape[1]
    etas = _powers(z3, 5)
    mc = m.clamp(max=2.0)
    phi2 = torch.zeros_like(T)
    phi3 = torch.zeros_like(T)
    for i in range(n):
        for j in range(i, n):
            sij3 = (0.5 * (sigma[:, i] + sigma[:, j])) ** 3
            mij = torch.sqrt(mc[:, i] * mc[:, j])
            w1 = (mij - 1.0) / mij
            w2 = w1 * (mij - 2.0) / mij
            eij = torch.sqrt(eps[:, i] * eps[:, j]) / T
            pair = _series(_AD, w1, w2, etas) + eij * _series(_BD, w1, w2, etas)
            mult = 1.0 if i == j else 2.0
            phi2 = phi2 - rho[:, i] * rho[:, j] * mu2_term[:, i] * mu2_term[:, j] * pair / sij3 * mult
            for k in range(j, n):
                sij = 0.5 * (sigma[:, i] + sigma[:, j])
                sik = 0.5 * (sigma[:, i] + sigma[:, k])
                sjk = 0.5 * (sigma[:, j] + sigma[:, k])
                mijk = (mc[:, i] * mc[:, j] * mc[:, k]) ** (1.0 / 3.0)
                v1 = (mijk - 1.0) / mijk
                v2 = v1 * (mijk - 2.0) / mijk
                mult3 = {1: 1.0, 2: 3.0, 3: 6.0}[len({i, j, k})]
                phi3 = phi3 - (
                    rho[:, i] * rho[:, j] * rho[:, k] * mu2_term[:, i] * mu2_term[:, j] * mu2_term[:, k]
                    * _series(_CD, v1, v2, etas) / (sij * sik * sjk) * mult3
                )
    phi2 = phi2 * PI
    phi3 = phi3 * (4.0 / 3.0 * PI * PI)
    dipolar = (mu2_term.detach() > 0).any(dim=1)
    den = torch.where(dipolar, phi2 - phi3, torch.ones_like(phi2))
    return torch.where(dipolar, phi2 * phi2 / den, torch.zeros_like(phi2))


def _strength(di, dj, si, sj, ki, kj, eps_aibj, T, z2, z3m1):
    """Association strength Delta^{AiBj} (pcsaft_mix.py:500-522, gc_pcsaft.py:545-564)."""
    k = di * dj / (di + dj) * z2 * z3m1
    return z3m1 * (k * (2.0 * k + 3.0) + 1.0) * (si * sj) ** 1.5 * torch.sqrt(ki * kj) * (torch.exp(eps_aibj / T) - 1.0)


def _newton_fixed_point(residual, x0, n_plain=60, n_diff=3, damp=True):
    """Solve residual(x)=0 for x [N,k] (k = 1 or 2) per state.

    The reference iterates Newton from 0.2 with the rule "negative iterate -> 0.2 * old"
    (pcsaft_mix.py:270-311, 363-385) and carries dual parts through the iteration.  Here the
    value is converged without autograd, then ``n_diff`` differentiable Newton steps are applied:
    each step from an exact root raises the order of exact derivatives (1, 3, 7, ...), which gives
    the same derivatives as implicit differentiation of the converged site fractions."""

    def step(x, create_graph):
        k = x.shape[1]
        xs = x.detach().clone().requires_grad_(True) if not create_graph else x
        f = residual(xs)
        cols = []
        for c in range(k):
            (g,) = torch.autograd.grad(f[:, c].sum(), xs, create_graph=create_graph, retain_graph=True)
            cols.append(g)
        jac = torch.stack(cols, dim=1)  # [N, k(eq), k(var)]
        if k == 1:
            dx = f / jac[:, 0, :]
        else:
            det = jac[:, 0, 0] * jac[:, 1, 1] - jac[:, 0, 1] * jac[:, 1, 0]
            dx0 = (jac[:, 1, 1] * f[:, 0] - jac[:, 0, 1] * f[:, 1]) / det
            dx1 = (-jac[:, 1, 0] * f[:, 0] + jac[:, 0, 0] * f[:, 1]) / det
            dx = torch.stack([dx0, dx1], dim=1)
        return xs, f, dx

    x = x0.detach().clone()
    with torch.enable_grad():
        for _ in range(n_plain):
            xs, f, dx = step(x, False)
            new = (xs - dx).detach()
            if damp:
                new = torch.where(new < 0, 0.2 * x, new)
            done = bool((f.detach().abs().max() < 1e-15) or (new - x).abs().max() < 1e-16)
            x = new
            if done:
                break
    if x.shape[0] == 0:
        return x
    needs_graph = torch.is_grad_enabled()
    if not needs_graph:
        return x
    for _ in range(n_diff):
        xs, f, dx = step(x if x.requires_grad else x.clone().requires_grad_(True), True)
        x = xs - dx
    return x


# --------------------------------------------------------------------------------------------
# binary mixture with k_ij                                              pcsaft_mix.py:12-154
# --------------------------------------------------------------------------------------------
def mix_helmholtz(params, kij, temperature, density):
    """a(T, rho_1, rho_2) [N] of a PC-SAFT mixture.  params [N,n,8]; kij [N,2] = (k_ij, eps_AiBj)
    (pcsaft_mix.py:13-29, 31-154).  Association categories follow :118-152."""
    m, sigma, eps, mu, kappa, eps_ab, na, nb = params.unbind(dim=-1)
    T, rho = temperature, density
    n = m.shape[1]
    Tn = T[:, None]
    d = sigma * (1.0 - 0.12 * torch.exp(-3.0 * eps / Tn))
    mr = m * rho
    z0 = PI / 6.0 * mr.sum(dim=1)
    z1 = PI / 6.0 * (mr * d).sum(dim=1)
    z2 = PI / 6.0 * (mr * d * d).sum(dim=1)
    z3 = PI / 6.0 * (mr * d * d * d).sum(dim=1)
    om = 1.0 - z3
    z3m1 = 1.0 / om

    hs = _hard_sphere(z0, z1, z2, z3)
    c = (z2 / (om * om))[:, None]
    g = z3m1[:, None] + 1.5 * d * c + 0.5 * d * d * c * c * om[:, None]
    hc = -(rho * (m - 1.0) * torch.log(g)).sum(dim=1)

    rho_tot = rho.sum(dim=1)
    mbar = (rho * m).sum(dim=1) / rho_tot
    r1 = torch.zeros_like(T)
    r2 = torch.zeros_like(T)
    for i in range(n):
        for j in range(n):
            eij = torch.sqrt(eps[:, i] * eps[:, j]) / T
            if i != j:
                eij = eij * (1.0 - kij[:, 0])
            term = rho[:, i] * rho[:, j] * m[:, i] * m[:, j] * eij * (0.5 * (sigma[:, i] + sigma[:, j])) ** 3
            r1 = r1 + term
            r2 = r2 + term * eij
    phi = hs + hc + _dispersion_tail(r1, r2, mbar, z3)

    mu2 = mu * mu / (m * sigma**3 * eps) * DIPOLE_SCALE
    phi = phi + _dipole_binary(m, sigma, eps, sigma**3 * eps * mu2 / Tn, T, rho, z3)

    # association categories (pcsaft_mix.py:118-152)
    n_assoc = torch.count_nonzero((na + nb).detach(), dim=1)
    n_self = torch.count_nonzero((na * nb).detach(), dim=1)
    if bool((n_assoc > 2).any()):
        raise Exception("Only up to two associating components are allowed!")
    sel_self = (n_assoc == 1) & (n_self == 1)
    sel_cross = (n_assoc == 2) & (n_self == 2)
    sel_ind = (n_assoc == 2) & (n_self == 1)
    if bool((sel_cross | sel_ind).any()) and n != 2:
        raise Exception("Cross/induced association is only implemented for binary mixtures!")

    pieces = torch.zeros_like(T)
    if bool(sel_self.any()):
        s = sel_self
        pieces = pieces.index_put((s,), _mix_self_assoc(kappa[s], eps_ab[s], na[s], nb[s], sigma[s], d[s], T[s], rho[s], z2[s], z3m1[s]))
    if bool(sel_cross.any()):
        s = sel_cross
        pieces = pieces.index_put((s,), _mix_cross_assoc(kappa[s], eps_ab[s], na[s], nb[s], sigma[s], d[s], T[s], rho[s], z2[s], z3m1[s], kij[s, 1]))
    if bool(sel_ind.any()):
        s = sel_ind
        pieces = pieces.index_put((s,), _mix_induced_assoc(kappa[s], eps_ab[s], na[s], nb[s], sigma[s], d[s], T[s], rho[s], z2[s], z3m1[s]))
    return phi + pieces


def _mix_self_assoc(kappa, eps_ab, na, nb, sigma, d, T, rho, z2, z3m1):
    """One self-associating component (pcsaft_mix.py:210-239): closed-form site fractions."""
    kap = kappa.sum(dim=1)
    eab = eps_ab.sum(dim=1)
    sig = (na * sigma).sum(dim=1) / na.sum(dim=1)
    dd = (na * d).sum(dim=1) / na.sum(dim=1)
    delta = _strength(dd, dd, sig, sig, kap, kap, eab, T, z2, z3m1)
    ra = (na * rho).sum(dim=1)
    rb = (nb * rho).sum(dim=1)
    aux = 1.0 + (ra - rb) * delta
    root = torch.sqrt(aux * aux + 4.0 * rb * delta)
    xa = 2.0 / (root + 1.0 + (rb - ra) * delta)
    xb = 2.0 / (root + 1.0 + (ra - rb) * delta)
    return ra * _site_term(xa) + rb * _site_term(xb)


def _pair_strengths(kappa, eps_ab, sigma, d, T, z2, z3m1, eps_aibj):
    out = {}
    for i in range(2):
        for j in range(2):
            mean = 0.5 * (eps_ab[:, i] + eps_ab[:, j])
            if eps_aibj is not None and i != j:
                e = torch.where(eps_aibj != 0, eps_aibj, mean)
            else:
                e = mean
            out[i, j] = _strength(d[:, i], d[:, j], sigma[:, i], sigma[:, j], kappa[:, i], kappa[:, j], e, T, z2, z3m1)
    return out


def _mix_cross_assoc(kappa, eps_ab, na, nb, sigma, d, T, rho, z2, z3m1, eps_aibj):
    """Two self-associating components (pcsaft_mix.py:241-321): 2x2 Newton for X_A0, X_A1."""
    D = _pair_strengths(kappa, eps_ab, sigma, d, T, z2, z3m1, eps_aibj)
    ra = rho * na
    rb = rho * nb

    def xb_inv(x):
        return (
            1.0 + x[:, 0] * ra[:, 0] * D[0, 0] + x[:, 1] * ra[:, 1] * D[0, 1],
            1.0 + x[:, 0] * ra[:, 0] * D[1, 0] + x[:, 1] * ra[:, 1] * D[1, 1],
        )

    def residual(x):
        b0, b1 = xb_inv(x)
        f0 = x[:, 0] - 1.0 + x[:, 0] / b0 * rb[:, 0] * D[0, 0] + x[:, 0] / b1 * rb[:, 1] * D[0, 1]
        f1 = x[:, 1] - 1.0 + x[:, 1] / b0 * rb[:, 0] * D[1, 0] + x[:, 1] / b1 * rb[:, 1] * D[1, 1]
        return torch.stack([f0, f1], dim=1)

    x = _newton_fixed_point(residual, torch.full_like(rho, 0.2))
    b0, b1 = xb_inv(x)
    return (
        ra[:, 0] * _site_term(x[:, 0]) + ra[:, 1] * _site_term(x[:, 1])
        + rb[:, 0] * _site_term(1.0 / b0) + rb[:, 1] * _site_term(1.0 / b1)
    )


def _induced_solve(D, na, nb, rho):
    """Shared by PcSaftMix/GcPcSaftMix induced association (pcsaft_mix.py:323-393,
    gc_pcsaft.py:382-441): one unknown X_A, B-sites on both components."""
    d00, d01 = D[0, 0] * rho[:, 0], D[0, 1] * rho[:, 1]
    d10, d11 = D[1, 0] * rho[:, 0], D[1, 1] * rho[:, 1]
    na0, na1, nb0, nb1 = na[:, 0], na[:, 1], nb[:, 0], nb[:, 1]

    def binv(x):
        return 1.0 + x * (na0 * d00 + na1 * d01), 1.0 + x * (na0 * d10 + na1 * d11)

    def residual(xx):
        x = xx[:, 0]
        b0, b1 = binv(x)
        f0 = x * (b0 * b1 + nb0 * b1 * d00 + nb1 * b0 * d01) - b0 * b1
        f1 = x * (b0 * b1 + nb0 * b1 * d10 + nb1 * b0 * d11) - b0 * b1
        return (na0 * f0 + na1 * f1)[:, None]

    x = _newton_fixed_point(residual, torch.full_like(rho[:, :1], 0.2))[:, 0]
    b0, b1 = binv(x)
    return rho[:, 0] * (_site_term(x) * na0 + _site_term(1.0 / b0) * nb0) + rho[:, 1] * (
        _site_term(x) * na1 + _site_term(1.0 / b1) * nb1
    )


def _mix_induced_assoc(kappa, eps_ab, na, nb, sigma, d, T, rho, z2, z3m1):
    D = _pair_strengths(kappa, eps_ab, sigma, d, T, z2, z3m1, None)
    return _induced_solve(D, na, nb, rho)


def mixture_derivatives(helmholtz, temperature, density):
    """(a, p/kT, mu_i^res/kT, v_i) of a mixture model ``helmholtz(T, rho[N,n])``.

    Reference: pcsaft_mix.py:395-420 / gc_pcsaft.py:443-468 (hyper-dual numbers in (N_i, V)).
    Here from the gradient a_i and Hessian a_ij of a in the partial densities:
      mu_i = a_i;  p = rho - a + sum_i rho_i a_i;
      v_i = (1 + sum_j rho_j a_ij) / (rho + sum_ij rho_i rho_j a_ij)."""
    rho = density.detach().clone().requires_grad_(True)
    a = helmholtz(temperature, rho)
    g = _grad(a, rho)
    n = rho.shape[1]
    hess = torch.stack([_grad(g[:, i], rho) for i in range(n)], dim=1)  # [N,n,n]
    r = density
    p = r.sum(dim=1) - a + (r * g).sum(dim=1)
    hr = (hess * r[:, None, :]).sum(dim=2)  # sum_j rho_j a_ij
    v = (1.0 + hr) / (r.sum(dim=1) + (r * hr).sum(dim=1))[:, None]
    return a, p, g, v


def bubble_pressure_from_densities(helmholtz, temperature, rho_v, rho_l):
    """Newton-step bubble pressure [Pa] at fixed partial densities (pcsaft_mix.py:435-444)."""
    rv = rho_v.sum(dim=1)
    y = rho_v / rv[:, None]
    _, p_l, mu_l, v_il = mixture_derivatives(helmholtz, temperature, rho_l)
    a_v = helmholtz(temperature, rho_v) / rv
    v_l = (y * v_il).sum(dim=1)
    g_l = (y * (torch.log(rho_v / rho_l) - mu_l)).sum(dim=1)
    p = -(a_v + p_l * v_l + g_l - 1.0) / (1.0 / rv - v_l)
    return p * temperature * PA_PER_RED


def dew_pressure_from_densities(helmholtz, temperature, rho_v, rho_l):
    """Newton-step dew pressure [Pa] at fixed partial densities (pcsaft_mix.py:459-468)."""
    rl = rho_l.sum(dim=1)
    x = rho_l / rl[:, None]
    _, p_v, mu_v, v_iv = mixture_derivatives(helmholtz, temperature, rho_v)
    a_l = helmholtz(temperature, rho_l) / rl
    v_v = (x * v_iv).sum(dim=1)
    g_v = (x * (torch.log(rho_l / rho_v) - mu_v)).sum(dim=1)
    p = -(a_l + p_v * v_v + g_v - 1.0) / (1.0 / rl - v_v)
    return p * temperature * PA_PER_RED


# --------------------------------------------------------------------------------------------
# heterosegmented group-contribution PC-SAFT                            gc_pcsaft.py:13-441
# --------------------------------------------------------------------------------------------
class GcTables:
    """Dense per-state tables built the way gc_pcsaft.py:14-86 does (counts [N,2,S], bonds
    [N,2,S,S] stored at [max,min], k_ab [S,S] symmetric, molecule-level dipole/association
    parameters).  ``parameter`` = 8 tensors of shape [S]; they may require grad."""

    def __init__(self, segment_identifier, parameter, segment_lists, bond_lists, binary_segment_records, phi=None):
        m, sigma, eps, mu, kappa, eps_ab, na, nb = parameter
        S = len(segment_identifier)
        index = {s: i for i, s in enumerate(segment_identifier)}
        N = len(segment_lists)
        counts = torch.zeros((N, 2, S), dtype=torch.float64)
        bonds = torch.zeros((N, 2, S, S), dtype=torch.float64)
        for k, (segs, bnds) in enumerate(zip(segment_lists, bond_lists)):
            for c in range(2):
                for s in segs[c]:
                    counts[k, c, index[s]] += 1.0
                for i, j in bnds[c]:
                    a, b = index[segs[c][i]], index[segs[c][j]]
                    bonds[k, c, max(a, b), min(a, b)] += 1.0
        self.counts, self.bonds = counts, bonds
        self.m_seg = counts * m  # gc_pcsaft.py:55
        self.sigma, self.eps = sigma, eps
        self.phi = phi
        kab = torch.zeros((S, S), dtype=torch.float64)
        for s1, s2, k in binary_segment_records:  # gc_pcsaft.py:60-63 (k may be a 0-dim tensor)
            upd = torch.zeros((S, S), dtype=torch.float64)
            upd[index[s1], index[s2]] = 1.0
            upd[index[s2], index[s1]] = 1.0
            kab = kab * (1.0 - upd) + upd * k
        self.kab = kab
        # molecule-level dipole parameters (gc_pcsaft.py:66-73)
        self.m_mix = self.m_seg.sum(dim=2)
        self.sigma_mix = ((self.m_seg * sigma**3).sum(dim=2) / self.m_mix) ** (1.0 / 3.0)
        self.eps_mix = (self.m_seg * eps).sum(dim=2) / self.m_mix
        self.mu2 = (counts * mu**2).sum(dim=2) / self.m_mix * DIPOLE_SCALE
        # association (gc_pcsaft.py:76-86)
        is_assoc = counts * torch.sign((kappa * eps_ab).detach())
        if bool((is_assoc.sum(dim=2) > 1).any()):
            raise Exception("Only up to one associating segment per component is allowed!")
        self.sigma_assoc = (is_assoc * sigma).sum(dim=2)
        self.eps_assoc = (is_assoc * eps).sum(dim=2)
        self.kappa_ab = (counts * kappa).sum(dim=2)
        self.eps_ab = (counts * eps_ab).sum(dim=2)
        self.na = (counts * na).sum(dim=2)
        self.nb = (counts * nb).sum(dim=2)


def gc_helmholtz(tab, temperature, density):
    """a(T, rho_1, rho_2) [N] for GcTables ``tab`` (gc_pcsaft.py:116-253).  Only segment types
    present in the batch are visited -- identical value, see SURVEY.md section 8(a17)."""
    T, rho = temperature, density
    present = torch.nonzero(tab.counts.sum(dim=(0, 1)) > 0).flatten().tolist()
    sig = tab.sigma
    d = sig * (1.0 - 0.12 * torch.exp(-3.0 * tab.eps[None, :] / T[:, None]))  # [N,S]
    ms = tab.m_seg  # [N,2,S]

    def zeta(k):
        return PI / 6.0 * ((ms * d[:, None, :] ** k).sum(dim=2) * rho).sum(dim=1)

    z0, z1, z2, z3 = zeta(0), zeta(1), zeta(2), zeta(3)
    om = 1.0 - z3
    z3m1 = 1.0 / om
    hs = _hard_sphere(z0, z1, z2, z3)

    c = z2 / (om * om)
    hc = torch.zeros_like(T)
    for a in present:
        for b in present:
            if b > a or not bool((tab.bonds[:, :, a, b] != 0).any()):
                continue
            cdab = c * d[:, a] * d[:, b] / (d[:, a] + d[:, b])
            g = z3m1 + 3.0 * cdab + 2.0 * cdab * cdab * om
            hc = hc - (rho * tab.bonds[:, :, a, b]).sum(dim=1) * torch.log(g)

    rho_tot = rho.sum(dim=1)
    mbar = (rho * ms.sum(dim=2)).sum(dim=1) / rho_tot
    r1 = torch.zeros_like(T)
    r2 = torch.zeros_like(T)
    for i in range(2):
        for j in range(2):
            for a in present:
                for b in present:
                    e2 = tab.eps[a] * tab.eps[b]
                    if tab.phi is not None:
                        e2 = e2 * tab.phi[:, i] * tab.phi[:, j]
                    eab = torch.sqrt(e2 + torch.zeros_like(T)) / T
                    if i != j:
                        eab = eab * (1.0 - tab.kab[a, b])
                    term = rho[:, i] * rho[:, j] * ms[:, i, a] * ms[:, j, b] * eab * (0.5 * (sig[a] + sig[b])) ** 3
                    r1 = r1 + term
                    r2 = r2 + term * eab
    phi = hs + hc + _dispersion_tail(r1, r2, mbar, z3)

    # molecule-level dipoles: mu2_term = mu2 / T (gc_pcsaft.py:262)
    phi = phi + _dipole_binary(tab.m_mix, tab.sigma_mix, tab.eps_mix, tab.mu2 / T[:, None], T, rho, z3)

    ke = (tab.kappa_ab * tab.eps_ab).detach()
    n_assoc = torch.count_nonzero(ke, dim=1)
    n_self = torch.count_nonzero((tab.na * tab.nb).detach(), dim=1)
    if bool((n_assoc > 2).any()):
        raise Exception("Only up to two associating segments are allowed!")
    sel_self = (n_assoc == 1) & (n_self == 1)
    sel_cross = (n_assoc == 2) & (n_self == 2)
    sel_ind = (n_assoc == 2) & (n_self == 1)
    dd = tab.sigma_assoc * (1.0 - 0.12 * torch.exp(-3.0 * tab.eps_assoc / T[:, None]))

    pieces = torch.zeros_like(T)
    if bool(sel_self.any()):
        s = sel_self  # gc_pcsaft.py:309-330, nA = nB = 1 hard-wired
        sg = tab.sigma_assoc[s].sum(dim=1)
        ek = tab.eps_assoc[s].sum(dim=1)
        kap = tab.kappa_ab[s].sum(dim=1)
        eab = tab.eps_ab[s].sum(dim=1)
        d1 = sg * (1.0 - 0.12 * torch.exp(-3.0 * ek / T[s]))
        r = (torch.sign(ke[s]) * rho[s]).sum(dim=1)
        dr = _strength(d1, d1, sg, sg, kap, kap, eab, T[s], z2[s], z3m1[s]) * r
        xa = 2.0 / (torch.sqrt(1.0 + 4.0 * dr) + 1.0)
        pieces = pieces.index_put((s,), r * (2.0 * torch.log(xa) - xa + 1.0))
    if bool(sel_cross.any()):
        s = sel_cross  # gc_pcsaft.py:332-380, nA = nB = 1 hard-wired
        D = _pair_strengths(tab.kappa_ab[s], tab.eps_ab[s], tab.sigma_assoc[s], dd[s], T[s], z2[s], z3m1[s], None)
        r = rho[s]
        d00, d01, d10, d11 = D[0, 0] * r[:, 0], D[0, 1] * r[:, 1], D[1, 0] * r[:, 0], D[1, 1] * r[:, 1]

        def residual(x):
            f0 = x[:, 0] + x[:, 0] * x[:, 0] * d00 + x[:, 0] * x[:, 1] * d01 - 1.0
            f1 = x[:, 1] + x[:, 1] * x[:, 0] * d10 + x[:, 1] * x[:, 1] * d11 - 1.0
            return torch.stack([f0, f1], dim=1)

        x = _newton_fixed_point(residual, torch.full_like(r, 0.2), damp=False)
        f = lambda t: 2.0 * torch.log(t) - t + 1.0
        pieces = pieces.index_put((s,), r[:, 0] * f(x[:, 0]) + r[:, 1] * f(x[:, 1]))
    if bool(sel_ind.any()):
        s = sel_ind  # gc_pcsaft.py:382-441
        D = _pair_strengths(tab.kappa_ab[s], tab.eps_ab[s], tab.sigma_assoc[s], dd[s], T[s], z2[s], z3m1[s], None)
        pieces = pieces.index_put((s,), _induced_solve(D, tab.na[s], tab.nb[s], rho[s]))
    return phi + pieces
