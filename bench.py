"""bench.py -- headline benchmark of the batched PC-SAFT phase-equilibrium hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--n STATES]

Workload (BASELINE.json configs[1], "C2"): PcSaftPure vapor_pressure + liquid_density, forward and
parameter gradients, on N_STATES synthetic non-associating (params, T) pairs per GPU (SURVEY.md
section 8(d): m~U[1,4], sigma~U[3,4], eps~U[150,350], T = tau*T_c with tau~U[0.5,0.95] plus a 1 %
super-critical slice, p = max(logU[1e5,1e7] Pa, 1.05 p_sat)).  One "step" = one pass over the batch:
vapor_pressure fwd+bwd and liquid_density fwd+bwd through the public API.

Prints ONE JSON line (rank 0).  `value` = states (each = one vapor-pressure solve+gradient and one
liquid-density solve+gradient) per second over all GPUs with inputs resident in HBM; `e2e` = the same
with pinned HOST inputs, H2D/D2H copies inside the timed region.  The default run appends short legs of
the other BASELINE.json configs under `extras` (c3, c4, c5: own value, e2e, roofline, failed fraction;
under torchrun the c4 / c5 legs run the NCCL all-reduce of the shared-leaf gradient every step).
`--impl reference` times the reference CPU path (compiled per-state solver restating feos, OpenMP over
all host threads, + the reference's Newton-step gradient stage in torch on the CPU) on a bounded sample
of the same workload.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import tempfile
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "pcsaft_vapor_pressure_solves_plus_grads_per_s"
UNIT = "states/s"
F_D2 = 906.0  # algorithmic flops of one (a, p, dp/drho) evaluation, SURVEY.md section 8(d)
F_REAL = 305.0  # algorithmic flops of one real Helmholtz evaluation
L2_FLUSH_BYTES = 256 << 20


def make_workload(n, seed, device):
    g = torch.Generator().manual_seed(seed)
    r = lambda: torch.rand(n, generator=g, dtype=torch.float64)
    m, sigma, eps = 1.0 + 3.0 * r(), 3.0 + r(), 150.0 + 200.0 * r()
    tc = eps * (0.6380 + 0.7070 * m - 0.06418 * m * m)
    tau = 0.50 + 0.45 * r()
    n_sc = n // 100
    sc = torch.randperm(n, generator=g)[:n_sc]
    tau[sc] = 1.1 + 0.4 * torch.rand(n_sc, generator=g, dtype=torch.float64)
    params = torch.zeros((n, 8), dtype=torch.float64)
    params[:, 0], params[:, 1], params[:, 2] = m, sigma, eps
    T = tau * tc
    p = torch.exp(math.log(1e5) + r() * math.log(100.0))
    return params, T, p


def clocks_sampler(path, index):
    q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    try:
        return subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "200", "-i", str(index)],
                                stdout=open(path, "w"), stderr=subprocess.DEVNULL)
    except Exception:
        return None


def parse_clocks(path):
    sm, mx, reasons = [], [], set()
    try:
        for line in open(path):
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
    except OSError:
        pass
    if not sm:
        return {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
    sm.sort()
    return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


CPU_CHUNK = 1 << 17  # states per gradient-stage chunk of the CPU arm (bounds the autograd graph: ~1 GB)


def cpu_c2_pass(params, T, p):
    """One pass of the C2 workload on the host the way the reference runs it: solve stage = one compiled
    per-state solve per state on all cores (oracle/cpu_solver.cpp = what src/pcsaft.rs:82-129 asks of feos, OpenMP
    for rayon), gradient stage = the reference's Newton-step formulas under torch autograd on the CPU
    (oracle/eos.py = pcsaft_pure.py:196-199, :210-215).  Returns (grad, seconds in solve stage, in gradient stage)."""
    from oracle import cpu_solver, eos

    n = T.shape[0]
    t0 = time.perf_counter()
    rv, rl, nans = cpu_solver.pure_vle(params, T)
    t_solve = time.perf_counter() - t0
    t0 = time.perf_counter()
    grad = torch.zeros_like(params)
    psat = torch.zeros(n, dtype=torch.float64)
    for a in range(0, n, CPU_CHUNK):
        sl = slice(a, min(a + CPU_CHUNK, n))
        ok = ~nans[sl]
        x = params[sl][ok].clone().requires_grad_(True)
        vp = eos.pure_vapor_pressure_from_densities(x, T[sl][ok], rv[sl][ok], rl[sl][ok])
        vp.sum().backward()
        g = torch.zeros((ok.shape[0], 8), dtype=torch.float64)
        g[ok] = x.grad
        grad[sl] += g
        ps = torch.zeros(ok.shape[0], dtype=torch.float64)
        ps[ok] = vp.detach()
        psat[sl] = ps
    t_grad = time.perf_counter() - t0
    pr = torch.maximum(p, 1.05 * psat)
    t0 = time.perf_counter()
    rho, nans2 = cpu_solver.pure_liquid_density(params, T, pr)
    t_solve += time.perf_counter() - t0
    t0 = time.perf_counter()
    for a in range(0, n, CPU_CHUNK):
        sl = slice(a, min(a + CPU_CHUNK, n))
        ok = ~nans2[sl]
        x = params[sl][ok].clone().requires_grad_(True)
        val = eos.pure_liquid_density_from_density(x, T[sl][ok], pr[sl][ok], rho[sl][ok])
        val.sum().backward()
        g = torch.zeros((ok.shape[0], 8), dtype=torch.float64)
        g[ok] = x.grad
        grad[sl] += g
    t_grad += time.perf_counter() - t0
    return grad, t_solve, t_grad


def run_reference(args, rank, world, out):
    """Reference arm: the reference's CPU implementation of the path (compiled solver restating feos + the
    reference's Newton-step gradient stage under torch autograd), all host threads."""
    if rank != 0:
        return
    from oracle import cpu_solver

    n = args.ref_n
    threads = _use_all_host_threads()
    cpu_solver.load()
    params, T, p = make_workload(n, 1000, "cpu")
    for _ in range(args.warmup):
        cpu_c2_pass(params, T, p)
    t0 = time.perf_counter()
    ts = tg = 0.0
    for _ in range(args.steps):
        _, a, b = cpu_c2_pass(params, T, p)
        ts += a
        tg += b
    dt = time.perf_counter() - t0
    value = n * args.steps / dt
    sample = (f"{n} states per step of the C2 workload (seed 1000): compiled per-state solver (g++ -O3 -fopenmp, "
              f"{cpu_solver.threads()} OpenMP threads) + torch-CPU gradient stage ({threads} threads); "
              f"solve stage {ts / args.steps:.3f} s, gradient stage {tg / args.steps:.3f} s per step")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": C2.name, "states_per_step": n},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                         "solve_stage_states_per_s": n * args.steps / ts, "gradient_stage_states_per_s": n * args.steps / tg},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), file=out, flush=True)


def _bump_version(t):
    """What an optimizer step does to a leaf, without the update kernel: the tensor's version counter moves, so
    nothing keyed on (identity, version) of the parameters can be reused across steps for free."""
    inc = getattr(torch.autograd.graph, "increment_version", None)
    if inc is not None:
        inc(t)
    else:
        with torch.no_grad():
            t.add_(0.0)


class C2:
    """BASELINE.json configs[1]: PcSaftPure vapor_pressure + liquid_density, fwd + parameter gradients."""
    name = "C2: PcSaftPure vapor_pressure + liquid_density fwd+bwd, synthetic non-associating (params,T) pairs"
    dominant = "feos_dev_pure_vle_fwd"
    kernel = "pure_vle_fwd_kernel(+retry)"
    flops_per_eval = F_D2
    io_bytes_per_state = 64 + 8 + 16 + 8 + 1
    # dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture of the dominant kernel at 1 M
    # states, per state (profiles/README.md)
    traffic_per_state = 124.2  # dram__bytes_read.sum + dram__bytes_write.sum per launch / 1 M states (gathered rows)
    traffic_source = "profiles/r2h_c2_raw.csv"
    h2d_per_state = 64 + 8 + 8
    d2h_per_state = 64 + 2 * 8 + 2
    default_n = 1_000_000

    def __init__(self, n, rank, dev):
        from feos_torch_b200 import PcSaftPure

        self.PcSaftPure, self.n, self.dev = PcSaftPure, n, dev
        params_h, T_h, p_h = make_workload(n, 1000 + rank, "cpu")
        self.params_h, self.T_h, self.p_h = params_h.pin_memory(), T_h.pin_memory(), p_h.pin_memory()
        self.params = self.params_h.to(dev).requires_grad_(True)
        self.T = self.T_h.to(dev)
        # workload generation (untimed): p >= 1.05 p_sat for the converged states
        with torch.no_grad():
            nans0, vp0 = PcSaftPure(self.params.detach()).vapor_pressure(self.T)
            psat = torch.zeros(n, dtype=torch.float64, device=dev)
            psat[~nans0] = vp0
            self.p = torch.maximum(self.p_h.to(dev), 1.05 * psat)
            self.p_h.copy_(self.p)
        self.shared = []

    def step_device(self):
        self.params.grad = None
        _bump_version(self.params)  # as if an optimizer had updated the per-state parameters in place
        model = self.PcSaftPure(self.params)
        nans, vp = model.vapor_pressure(self.T)
        nans2, rho = model.liquid_density(self.T, self.p)
        (vp.sum() + rho.sum()).backward()
        return nans

    def step_e2e(self):
        x = self.params_h.requires_grad_(True)
        x.grad = None
        model = self.PcSaftPure(x)  # host tensors in: H2D inside
        nans, vp = model.vapor_pressure(self.T_h)
        nans2, rho = model.liquid_density(self.T_h, self.p_h)
        loss = vp.sum() + rho.sum()
        loss.backward()  # gradient lands on the host tensor: D2H inside
        return float(loss.detach()), x.grad

    def mean_evals(self, _lib):
        n = self.n
        evals = torch.zeros(n, dtype=torch.int32, device=self.dev)
        rho = torch.empty((n, 2), dtype=torch.float64, device=self.dev)
        failed = torch.empty(n, dtype=torch.uint8, device=self.dev)
        from feos_torch_b200.pcsaft_pure import _state_order

        order = _state_order(self.params.detach(), self.T, (self.params, self.T))  # same order (hence kernel variant) as the timed calls
        _lib.call("feos_dev_pure_vle_fwd", _lib.ptr(self.params.detach()), _lib.ptr(self.T), n, _lib.ptr(rho), None, None,
                  _lib.ptr(failed), _lib.ptr(evals), _lib.ptr(order), torch.cuda.current_stream().cuda_stream)
        return evals.double().mean().item()


def critical_temperature(PcSaftPure, params, t_lo, t_hi, iters=22):
    """Per-state T_c estimate by bisection on the converged/failed mask of the VLE kernel (workload
    generation only, untimed): the largest T in [t_lo, t_hi] with a vapour-liquid equilibrium."""
    model = PcSaftPure(params)
    lo, hi = t_lo.clone(), t_hi.clone()
    for _ in range(iters):
        mid = 0.5 * (lo + hi)
        failed, _ = model.vapor_pressure(mid)
        lo = torch.where(failed, lo, mid)
        hi = torch.where(failed, mid, hi)
    return lo


class C3:
    """BASELINE.json configs[2]: PcSaftPure equilibrium_liquid_density with 2B-associating parameters
    (kappa_AB ~ logU[0.003, 0.1], eps_AB/k ~ U[1500, 3000] K, nA = nB = 1), T = tau T_c(theta), tau ~ U[0.55, 0.95]
    with the per-state T_c from `critical_temperature`."""
    name = "C3: PcSaftPure equilibrium_liquid_density fwd+bwd, synthetic 2B-associating (params,T) pairs"
    dominant = "feos_dev_pure_vle_fwd"
    kernel = "pure_vle_fwd_kernel(+retry)"
    flops_per_eval = F_D2
    io_bytes_per_state = 64 + 8 + 16 + 8 + 1
    traffic_per_state = 384.3  # 3.0 GB read + 0.85 GB written per 10 M states (gathered rows + spills of the association path)
    traffic_source = "profiles/r2h_c3_raw.csv"
    h2d_per_state = 64 + 8
    d2h_per_state = 64 + 8 + 1
    default_n = 10_000_000

    def __init__(self, n, rank, dev):
        from feos_torch_b200 import PcSaftPure

        self.PcSaftPure, self.n, self.dev = PcSaftPure, n, dev
        g = torch.Generator().manual_seed(4000 + rank)
        r = lambda: torch.rand(n, generator=g, dtype=torch.float64)
        params = torch.zeros((n, 8), dtype=torch.float64)
        params[:, 0], params[:, 1], params[:, 2] = 1.0 + 3.0 * r(), 3.0 + r(), 150.0 + 200.0 * r()
        params[:, 4] = torch.exp(math.log(0.003) + r() * math.log(0.1 / 0.003))
        params[:, 5] = 1500.0 + 1500.0 * r()
        params[:, 6] = 1.0
        params[:, 7] = 1.0
        tau = 0.55 + 0.40 * r()
        self.params_h = params.pin_memory()
        self.params = self.params_h.to(dev).requires_grad_(True)
        with torch.no_grad():
            m, eps = self.params[:, 0].detach(), self.params[:, 2].detach()
            tc0 = eps * (0.6380 + 0.7070 * m - 0.06418 * m * m)  # non-associating T_c: a lower bound
            tc = critical_temperature(PcSaftPure, self.params.detach(), 0.9 * tc0, 6.0 * tc0)
            self.T = (tau.to(dev) * tc).contiguous()
        self.T_h = self.T.cpu().pin_memory()
        self.shared = []

    def step_device(self):
        self.params.grad = None
        _bump_version(self.params)
        nans, rho = self.PcSaftPure(self.params).equilibrium_liquid_density(self.T)
        rho.sum().backward()
        return nans

    def step_e2e(self):
        x = self.params_h.requires_grad_(True)
        x.grad = None
        nans, rho = self.PcSaftPure(x).equilibrium_liquid_density(self.T_h)
        loss = rho.sum()
        loss.backward()
        return float(loss.detach()), x.grad

    def mean_evals(self, _lib):
        return C2.mean_evals(self, _lib)


class C4:
    """BASELINE.json configs[3]: PcSaftMix bubble + dew point of ONE shared pair with a shared k_ij leaf,
    fit loss sum (p - p_target)^2, k_ij gradient summed over the GPUs by one NCCL all-reduce."""
    name = "C4: PcSaftMix binary bubble_point + dew_point fwd+bwd with shared k_ij, fit-loss gradient all-reduced"
    dominant = "feos_dev_mix_point_fwd"
    kernel = "mix_point_fwd_kernel(+retry)"
    flops_per_eval = 3485.0  # binary non-associating hyper-dual evaluation, SURVEY.md section 8(d)
    io_bytes_per_state = 128 + 16 + 8 + 8 + 32 + 8 + 1
    traffic_per_state = 44.7  # mean of the bubble and dew launch
    traffic_source = "profiles/r2h_c4_raw.csv"
    h2d_per_state = 8 + 8
    d2h_per_state = 2
    default_n = 1_000_000

    def __init__(self, n, rank, dev):
        from feos_torch_b200 import PcSaftMix

        self.PcSaftMix, self.n, self.dev = PcSaftMix, n, dev
        g = torch.Generator().manual_seed(2000 + rank)
        pair = torch.tensor([[1, 3.5, 150, 0, 0, 0, 0, 0], [1, 3.5, 200, 0, 0, 0, 0, 0]], dtype=torch.float64)
        self.pair = pair.to(dev)
        self.kij = torch.tensor([-0.1, 0.0], dtype=torch.float64, device=dev, requires_grad=True)
        self.T_h = ((0.6 + 0.25 * torch.rand(n, generator=g, dtype=torch.float64)) * 150 * 1.2757).pin_memory()
        self.z_h = (0.05 + 0.9 * torch.rand(n, generator=g, dtype=torch.float64)).pin_memory()
        self.T, self.z = self.T_h.to(dev), self.z_h.to(dev)
        self.p0 = torch.full((n,), 1e5, dtype=torch.float64, device=dev)
        with torch.no_grad():
            m = PcSaftMix(self.pair.expand(n, 2, 8), self.kij.detach().expand(n, 2))
            self.tb = m.bubble_point(self.T, self.z, self.p0)[0] * 1.01
            self.td = m.dew_point(self.T, self.z, self.p0)[0] * 1.01
        self.shared = [self.kij]

    def _loss(self, T, z):
        n = self.n
        model = self.PcSaftMix(self.pair.expand(n, 2, 8), self.kij.expand(n, 2))
        pb, nb = model.bubble_point(T, z, self.p0)
        pd, nd = model.dew_point(T, z, self.p0)
        return ((pb - self.tb) ** 2).sum() + ((pd - self.td) ** 2).sum(), nb

    def step_device(self):
        from feos_torch_b200 import distributed as fd

        self.kij.grad = None
        loss, nans = self._loss(self.T, self.z)
        loss.backward()
        fd.all_reduce_gradients(self.shared)
        return nans

    def step_e2e(self):
        from feos_torch_b200 import distributed as fd

        self.kij.grad = None
        loss, nans = self._loss(self.T_h.to(self.dev, non_blocking=True), self.z_h.to(self.dev, non_blocking=True))
        loss.backward()
        fd.all_reduce_gradients(self.shared)
        return float(loss.detach()), self.kij.grad.cpu()

    def mean_evals(self, _lib):
        n = self.n
        evals = torch.zeros(n, dtype=torch.int32, device=self.dev)
        rho = torch.empty((n, 4), dtype=torch.float64, device=self.dev)
        failed = torch.empty(n, dtype=torch.uint8, device=self.dev)
        P = self.pair.expand(n, 2, 8).contiguous()
        K = self.kij.detach().expand(n, 2).contiguous()
        total = 0.0
        for bubble in (1, 0):  # the timed step makes one bubble and one dew call: mean over both
            evals.zero_()
            _lib.call("feos_dev_mix_point_fwd", _lib.ptr(P), _lib.ptr(K), _lib.ptr(self.T), _lib.ptr(self.z), None, n, bubble, 0,
                      _lib.ptr(rho), None, _lib.ptr(failed), _lib.ptr(evals), None, torch.cuda.current_stream().cuda_stream)
            total += evals.double().mean().item()
        return 0.5 * total


class C5:
    """BASELINE.json configs[4]: GcPcSaftMix bubble + dew point, group parameters of the reference's
    sauer2014_hetero.json (tests/golden/gc.json), 1 M states over the five non-associating molecule pairs of
    tests/test_gc_pcsaft.py:17-29, shared k_ab leaf with its gradient all-reduced."""
    name = "C5: GcPcSaftMix bubble_point + dew_point fwd+bwd, sauer2014 groups, 5 unique molecule pairs, shared k_ab"
    dominant = "feos_dev_gc_point_fwd"
    kernel = "gc_point_fwd_kernel(+retry)"
    flops_per_eval = 4687.0  # GC compact hyper-dual evaluation, SURVEY.md section 8(d)
    io_bytes_per_state = 4 + 16 + 8 + 8 + 32 + 8 + 1
    traffic_per_state = 127.3  # mean of the bubble and dew launch
    traffic_source = "profiles/r2h_c5_raw.csv"
    h2d_per_state = 8 + 8
    d2h_per_state = 2
    default_n = 1_000_000

    def __init__(self, n, rank, dev):
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import cases
        from feos_torch_b200 import GcPcSaftMix

        table = json.load(open(os.path.join(ROOT, "tests", "golden", "gc.json")))["segments"]
        self.ident = list(table.keys())
        self.par = tuple(torch.tensor([table[i][k] for i in self.ident], dtype=torch.float64, device=dev) for k in range(8))
        g = torch.Generator().manual_seed(3000 + rank)
        which = torch.randint(0, 5, (n,), generator=g).tolist()
        pairs = [(cases.GC_SEGMENTS[i], cases.GC_BONDS[i]) for i in range(5)]
        self.segs = [pairs[w][0] for w in which]
        self.bonds = [pairs[w][1] for w in which]
        self.kab = torch.tensor(0.03, dtype=torch.float64, device=dev, requires_grad=True)
        self.phi = torch.tensor([cases.GC_PHI], dtype=torch.float64, device=dev).expand(n, 2).contiguous()
        self.T_h = (150.0 + 100.0 * torch.rand(n, generator=g, dtype=torch.float64)).pin_memory()
        self.z_h = (0.05 + 0.9 * torch.rand(n, generator=g, dtype=torch.float64)).pin_memory()
        self.T, self.z = self.T_h.to(dev), self.z_h.to(dev)
        self.p0 = torch.full((n,), 1e5, dtype=torch.float64, device=dev)
        self.GcPcSaftMix, self.n, self.dev = GcPcSaftMix, n, dev
        self.model = self._model()  # the pair table is built once per fit, not per step
        self.shared = [self.kab]

    def _model(self):
        return self.GcPcSaftMix(self.ident, self.par, self.segs, self.bonds, [("CH3", "CH=O", self.kab), (">CH", "HCOO", -0.01)], self.phi)

    def _loss(self, T, z):
        m = self.model
        m._kab = torch.stack([self.kab, torch.tensor(-0.01, dtype=torch.float64, device=self.dev)])  # fresh graph each step
        pb, nb = m.bubble_point(T, z, self.p0)
        pd, nd = m.dew_point(T, z, self.p0)
        return pb.sum() + pd.sum(), nb

    def step_device(self):
        from feos_torch_b200 import distributed as fd

        self.kab.grad = None
        loss, nans = self._loss(self.T, self.z)
        loss.backward()
        fd.all_reduce_gradients(self.shared)
        return nans

    def step_e2e(self):
        from feos_torch_b200 import distributed as fd

        self.kab.grad = None
        loss, nans = self._loss(self.T_h.to(self.dev, non_blocking=True), self.z_h.to(self.dev, non_blocking=True))
        loss.backward()
        fd.all_reduce_gradients(self.shared)
        return float(loss.detach()), self.kab.grad.cpu()

    def mean_evals(self, _lib):
        n = self.n
        m = self.model
        evals = torch.zeros(n, dtype=torch.int32, device=self.dev)
        rho = torch.empty((n, 4), dtype=torch.float64, device=self.dev)
        failed = torch.empty(n, dtype=torch.uint8, device=self.dev)
        total = 0.0
        for bubble in (1, 0):
            evals.zero_()
            _lib.call("feos_dev_gc_point_fwd", _lib.ptr(m._pairs), m._n_pairs, _lib.ptr(m._seg.detach()), m._seg.shape[0],
                      _lib.ptr(m._kab.detach()), m._kab.shape[0], _lib.ptr(m._pair_index), _lib.ptr(m._phi), _lib.ptr(self.T),
                      _lib.ptr(self.z), None, n, bubble, 0, _lib.ptr(rho), None, _lib.ptr(failed), _lib.ptr(evals), None,
                      torch.cuda.current_stream().cuda_stream)
            total += evals.double().mean().item()
        return 0.5 * total


WORKLOADS = {"c2": C2, "c3": C3, "c4": C4, "c5": C5}


def _quiet_stdout():
    """Libraries (NCCL with NCCL_DEBUG=VERSION/INFO) print to fd 1; the contract is ONE JSON line on stdout.
    Route fd 1 to stderr for the run and return a writer onto the real stdout for that line."""
    sys.stdout.flush()
    real = os.dup(1)
    os.dup2(2, 1)
    return os.fdopen(real, "w")


def main():
    out = _quiet_stdout()
    try:
        _main(out)
    finally:
        out.flush()


METRICS = {"c2": METRIC, "c3": "pcsaft_equilibrium_liquid_density_solves_plus_grads_per_s",
           "c4": "pcsaft_binary_bubble_plus_dew_solves_plus_grads_per_s",
           "c5": "gc_pcsaft_binary_bubble_plus_dew_solves_plus_grads_per_s"}

# Hardware truth beside the algorithmic-flop fraction (profiles/README.md, ncu --set full of the dominant kernel at
# the bench size): FP64-pipe utilisation, lanes per instruction, and executed FP64 flops (2 DFMA + DADD + DMUL per
# cycle x cycles / duration) over the measured peak.  The closed-form evaluations execute fewer flops than the
# algorithmic count of SURVEY.md section 8(d) credits them with, so `frac` is NOT pipe utilisation.
NCU = {
    "c2": {"fp64_pipe_pct": 64.0, "lanes_per_inst": 29.5, "executed_flops_per_eval": 408.0, "source": "profiles/r2h_c2_raw.csv"},
    "c3": {"fp64_pipe_pct": 64.3, "lanes_per_inst": 29.3, "executed_flops_per_eval": 730.0, "source": "profiles/r2h_c3_raw.csv"},
    "c4": {"fp64_pipe_pct": 62.4, "lanes_per_inst": 31.9, "executed_flops_per_eval": 1420.0, "source": "profiles/r2h_c4_raw.csv"},
    "c5": {"fp64_pipe_pct": 52.6, "lanes_per_inst": 31.0, "executed_flops_per_eval": 1776.0, "source": "profiles/r2h_c5_raw.csv"},
}


def measure(name, n, steps, warmup, rank, world, dev, flush, peak):
    """Build workload `name` with n states per GPU, time `steps` device-resident and `steps` end-to-end steps
    (CUDA events, L2 flushed between iterations, max over ranks) and return the fields of the JSON line."""
    import torch.distributed as dist

    from feos_torch_b200 import _lib

    wl = WORKLOADS[name](n, rank, dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, collect):
        import gc

        for _ in range(warmup):
            fn()
        # no Python garbage collection inside the timed steps: a generation-2 pass on ONE of N ranks costs milliseconds
        # and, with the time taken as the max over ranks, lands on the N-GPU number (measured at 8 GPUs)
        gc.collect()
        if fn.__name__ == "step_device":  # (the end-to-end steps are 4-90 ms long and hold GB of pinned buffers: left alone)
            gc.disable()
        barrier()
        total_ms = 0.0
        if collect:
            _lib.TIMINGS = {}
        profiled = os.environ.get("FEOS_BENCH_CUDA_PROFILER") == fn.__name__
        if profiled:  # `ncu --profile-from-start off`: only the timed steps of this leg (not the workload generation)
            torch.cuda.cudart().cudaProfilerStart()
        for _ in range(steps):
            flush.zero_()  # L2 flush between timed iterations (inputs are smaller than L2)
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            e1.synchronize()
            total_ms += e0.elapsed_time(e1)
        if profiled:
            torch.cuda.cudart().cudaProfilerStop()
        gc.enable()
        barrier()
        t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    launches0 = _lib.launch_count()
    ms_dev = timed(wl.step_device, False)  # the headline: no per-call events inside the timed steps
    launches = (_lib.launch_count() - launches0) * steps // (steps + warmup)
    # kernel shares: the same steps once more with CUDA events around every library call (costs ~10 us per call,
    # which is why the headline pass above runs without them)
    ms_dev_events = timed(wl.step_device, True)
    timings = _lib.TIMINGS
    _lib.TIMINGS = None
    ms_e2e = timed(wl.step_e2e, False)
    allreduce_us = None
    if wl.shared and world > 1:
        # the one collective of the path: sum of the shared-leaf gradient (a handful of doubles) over the ranks
        from feos_torch_b200 import distributed as fd

        for _ in range(3):
            fd.all_reduce_gradients(wl.shared)
        barrier()
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            fd.all_reduce_gradients(wl.shared)
        e1.record()
        e1.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) * 1e3 / 20], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        allreduce_us = t.item()

    # failed-state accounting + evaluation counts for the roofline (untimed)
    nans = wl.step_device()
    torch.cuda.synchronize()
    failed_frac = nans.double().mean().item()
    mean_evals = wl.mean_evals(_lib)
    if rank != 0:
        return None

    kern_ms = {k: sum(a.elapsed_time(b) for a, b in v) / steps for k, v in timings.items()}
    step_ms = ms_dev / steps
    step_ms_events = ms_dev_events / steps
    dom = wl.dominant
    dom_calls = len(timings[dom]) // steps
    flops_per_launch = n * mean_evals * wl.flops_per_eval
    dom_ms = kern_ms[dom] / dom_calls
    achieved = flops_per_launch / (dom_ms * 1e-3) / 1e12
    spec = 148 * 64 * 2 * 1.965e9 / 1e12
    io_bytes = n * wl.io_bytes_per_state
    solver_ms = sum(v for k, v in kern_ms.items() if k.startswith("feos_dev_") and not k.endswith("_order"))
    roofline = {
        "bound": "fp64", "kernel": wl.kernel, "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
        "frac": achieved / peak if peak else None,
        "traffic": wl.traffic_per_state * n if wl.traffic_per_state else None, "traffic_source": wl.traffic_source,
        "peak_source": "DFMA-chain microbenchmark in this run (MEASURED_PEAKS.json has no FP64 figure); spec 148 SM x 64 lanes x 2 x 1.965 GHz = %.1f" % spec,
        "flops_per_launch": flops_per_launch, "mean_evals_per_state": mean_evals, "flops_per_eval": wl.flops_per_eval,
        "kernel_ms": dom_ms, "share_of_step": kern_ms[dom] / step_ms_events, "solver_kernels_share_of_step": solver_ms / step_ms_events,
        "ms_per_step_with_call_events": step_ms_events,
        "ncu": NCU.get(name),
        "executed": _executed(NCU.get(name), n, mean_evals, dom_ms, peak),
        "note": "achieved / frac follow SURVEY.md section 8(d): ALGORITHMIC flops of the reference's (hyper-)dual evaluation x "
                "evaluations / kernel time; the kernels evaluate the same quantities in closed form with 2.2-2.5x fewer "
                "executed flops, so frac can exceed the pipe utilisation (and 1): `executed` = flops the kernel issues "
                "(ncu: 2 DFMA + DMUL + DADD) / time over the same peak, `ncu.fp64_pipe_pct` = FP64 pipe utilisation",
        "hbm": {"algorithmic_bytes": io_bytes, "achieved_gbs": io_bytes / (dom_ms * 1e-3) / 1e9, "peak_gbs": _peaks().get("hbm_gbs")},
        "kernel_ms_per_step": kern_ms,
    }
    par = f"batch sharded over {world} GPU(s)"
    if wl.shared:
        par += ", one NCCL all-reduce of the shared-leaf gradient per step" if world > 1 else ", shared-leaf gradient reduced in the backward kernel (all-reduce at N > 1)"
    else:
        par += ", no data-path collective"
    return {
        "metric": METRICS[name], "value": n * world * steps / (ms_dev * 1e-3), "unit": UNIT, "n_gpus": world,
        "steps": steps, "warmup": warmup, "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl.name, "states_per_gpu_per_step": n, "parallelism": par,
                   "l2": "flushed (256 MiB write) between timed iterations", "failed_fraction": failed_frac,
                   "state_order": "pure: recomputed every %d-th parameter version (parameters' version bumped every step); "
                                  "binary / GC: keyed on (T, x), computed once per data set" % _lib.ORDER_REFRESH},
        "e2e": {"value": n * world * steps / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": n * wl.h2d_per_state * world,
                "d2h_bytes_per_step": (n * wl.d2h_per_state + 8) * world, "ms_per_step": ms_e2e / steps},
        "gpu_launches": launches,
        "roofline": roofline,
        "allreduce_us_per_step": allreduce_us,
    }


def _executed(ncu, n, mean_evals, kernel_ms, peak):
    """Executed-flop rate of the dominant kernel: per-evaluation executed FP64 flops from the ncu capture named in
    `ncu["source"]` (profiles/flops_from_ncu.py) x the evaluations of THIS run / the kernel time of THIS run."""
    if not ncu or not ncu.get("executed_flops_per_eval"):
        return None
    tf = n * mean_evals * ncu["executed_flops_per_eval"] / (kernel_ms * 1e-3) / 1e12
    return {"tflops": tf, "frac_of_peak": tf / peak if peak else None, "flops_per_eval": ncu["executed_flops_per_eval"]}


def _main(out):
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS), help="c2 = headline (BASELINE.json metric)")
    ap.add_argument("--n", type=int, default=0, help="states per GPU per step (default: the workload's BASELINE size)")
    ap.add_argument("--ref-n", type=int, default=400_000, help="states per step of the reference arm (about 2 s per step)")
    ap.add_argument("--cpu-n", type=int, default=1_000_000, help="states of the cpu_baseline sample (3 passes, about 15 s)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--extras", default=None, help="comma-separated workloads measured in short legs under `extras` "
                    "(default: c3,c4,c5 with the headline workload, none otherwise; 'none' switches them off)")
    ap.add_argument("--extra-steps", type=int, default=10)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else max(args.warmup, 1)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world, out)
        return

    import torch.distributed as dist

    if world > 1 and os.environ.get("FEOS_BENCH_PIN", "1") != "0":
        # before any pinned host memory is allocated (first touch decides the NUMA node)
        _pin_to_gpu_cores(local, world, partition=os.environ.get("FEOS_BENCH_PIN", "1") == "2")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    # host-side reductions of the end-to-end leg (loss = sum over the returned host tensors): 4 threads per rank.
    # With all 16 vCPUs of the box in the OpenMP team the following device->host copies into the recycled pinned
    # buffers run at a fifth of PCIe speed (profiles/README.md, round 2: 763 vs 160 us per 8 MB); torchrun's default
    # of 1 thread makes the sums themselves the bottleneck.
    try:
        torch.set_num_threads(max(1, min(4, len(os.sched_getaffinity(0)) // max(1, world))))
    except AttributeError:
        torch.set_num_threads(4)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    from feos_torch_b200 import _lib

    _lib.load()
    flush = torch.empty(L2_FLUSH_BYTES // 8, dtype=torch.float64, device=dev)
    peak = _lib.measure_fp64_peak()

    clk_file = tempfile.NamedTemporaryFile(prefix="clocks_", suffix=".csv", delete=False).name
    sampler = clocks_sampler(clk_file, local) if rank == 0 else None
    n = args.n or WORKLOADS[args.workload].default_n
    line = measure(args.workload, n, args.steps, args.warmup, rank, world, dev, flush, peak)
    if args.extras is None:
        extras = ["c3", "c4", "c5"] if args.workload == "c2" and not args.n else []
    else:
        extras = [e for e in args.extras.split(",") if e and e != "none"]
    extra_lines = {}
    for name in extras:
        torch.cuda.empty_cache()
        r = measure(name, WORKLOADS[name].default_n, args.extra_steps, 3, rank, world, dev, flush, peak)
        if r is not None:
            extra_lines[name] = {k: r[k] for k in ("metric", "value", "unit", "steps", "ms_per_step", "config", "e2e", "gpu_launches", "roofline", "allreduce_us_per_step")}
    if sampler is not None:
        sampler.terminate()
        sampler.wait()
    if rank == 0:
        line["clocks"] = parse_clocks(clk_file)
        line["cpu_baseline"] = None
        if not args.no_cpu_baseline and args.workload == "c2":
            line["cpu_baseline"] = cpu_baseline(args.cpu_n)
        if extra_lines:
            line["extras"] = extra_lines
        print(json.dumps(line), file=out, flush=True)
    if world > 1:
        dist.destroy_process_group()


def _pin_to_gpu_cores(local, world=1, partition=False):
    """One process per GPU on a multi-socket box: run this rank (and allocate its pinned staging buffers) on the CPU
    cores the driver reports as local to its GPU (NVML cpu affinity = the cores of the GPU's NUMA node).
    partition=True (FEOS_BENCH_PIN=2) additionally gives every local rank its own contiguous share of that set."""
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local)
        ncpu = os.cpu_count() or 64
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (word >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus and partition and world > 1 and len(cpus) >= 2 * world:
            ordered = sorted(cpus)
            share = len(ordered) // world
            cpus = set(ordered[local * share:(local + 1) * share])
        if cpus:
            os.sched_setaffinity(0, cpus)
    except Exception as e:  # not fatal: the run is merely not pinned
        print(f"[bench] rank-to-core pinning skipped: {e}", file=sys.stderr)


def _use_all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1; the CPU arm is meant to use every host core it may run on."""
    try:
        cores = len(os.sched_getaffinity(0))
    except AttributeError:
        cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    return torch.get_num_threads()


def _peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


def cpu_baseline(n):
    """The reference CPU path on this box's host cores (rank 0, N = 1): compiled per-state solver + torch-CPU
    gradient stage on the first n states of the C2 workload; best of 3 passes after one warm-up pass of n / 4."""
    from oracle import cpu_solver

    cpu_solver.load()
    params, T, p = make_workload(n, 1000, "cpu")
    threads = _use_all_host_threads()
    k = max(1, n // 4)
    cpu_c2_pass(params[:k], T[:k], p[:k])
    best = None
    for _ in range(3):
        t0 = time.perf_counter()
        _, ts, tg = cpu_c2_pass(params, T, p)
        dt = time.perf_counter() - t0
        if best is None or dt < best[0]:
            best = (dt, ts, tg)
    dt, ts, tg = best
    return {"value": n / dt, "unit": UNIT, "cores": threads, "kind": "port",
            "solve_stage_states_per_s": n / ts, "gradient_stage_states_per_s": n / tg,
            "sample": f"first {n} states of the C2 workload, best of 3 passes: compiled per-state solver restating feos "
                      f"(oracle/cpu_solver.cpp, g++ -O3 -fopenmp, {cpu_solver.threads()} threads) {ts:.2f} s + the reference's "
                      f"Newton-step gradient stage under torch autograd on the CPU ({threads} threads) {tg:.2f} s"}


if __name__ == "__main__":
    main()
