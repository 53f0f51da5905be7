"""bench.py -- headline benchmark of the batched PC-SAFT phase-equilibrium hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--n STATES]

Workload (BASELINE.json configs[1], "C2"): PcSaftPure vapor_pressure + liquid_density, forward and
parameter gradients, on N_STATES synthetic non-associating (params, T) pairs per GPU (SURVEY.md
section 8(d): m~U[1,4], sigma~U[3,4], eps~U[150,350], T = tau*T_c with tau~U[0.5,0.95] plus a 1 %
super-critical slice, p = max(logU[1e5,1e7] Pa, 1.05 p_sat)).  One "step" = one pass over the batch:
vapor_pressure fwd+bwd and liquid_density fwd+bwd through the public API.

Prints ONE JSON line (rank 0).  `value` = states (each = one vapor-pressure solve+gradient and one
liquid-density solve+gradient) per second over all GPUs with inputs resident in HBM; `e2e` = the same
with pinned HOST inputs, H2D/D2H copies inside the timed region.  `--impl reference` times the
reference CPU path (oracle port, all host threads) on a bounded sample of the same workload.
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


def run_reference(args, rank, world):
    """Reference arm: the reference's CPU implementation of the path (oracle port: CPU solver restating
    feos + the reference's Newton-step gradient stage under torch autograd), all host threads."""
    if rank != 0:
        return
    from oracle import models as om

    n = args.ref_n
    threads = torch.get_num_threads()
    params, T, p = make_workload(n, 1000, "cpu")

    def step():
        x = params.clone().requires_grad_(True)
        model = om.PcSaftPure(x)
        nans, vp = model.vapor_pressure(T)
        psat = torch.zeros(n, dtype=torch.float64)
        psat[~nans] = vp.detach()
        pr = torch.maximum(p, 1.05 * psat)
        nans2, rho = model.liquid_density(T, pr)
        (vp.sum() + rho.sum()).backward()
        return x.grad

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    value = n * args.steps / dt
    sample = f"{n} states per step of the C2 workload (seed 1000), oracle port on {threads} torch CPU threads"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "C2: PcSaftPure vapor_pressure + liquid_density fwd+bwd, synthetic non-associating (params,T) pairs",
                   "states_per_step": n},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n", type=int, default=1_000_000, help="states per GPU per step")
    ap.add_argument("--ref-n", type=int, default=20_000, help="states per step of the reference arm / cpu_baseline")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else max(args.warmup, 1)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch.distributed as dist

    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    from feos_torch_b200 import PcSaftPure, _lib

    _lib.load()
    n = args.n
    params_h, T_h, p_h = make_workload(n, 1000 + rank, "cpu")
    params_h, T_h, p_h = params_h.pin_memory(), T_h.pin_memory(), p_h.pin_memory()
    params = params_h.to(dev).requires_grad_(True)
    T = T_h.to(dev)
    # workload generation (untimed): p >= 1.05 p_sat for the converged states
    with torch.no_grad():
        nans0, vp0 = PcSaftPure(params.detach()).vapor_pressure(T)
        psat = torch.zeros(n, dtype=torch.float64, device=dev)
        psat[~nans0] = vp0
        p = torch.maximum(p_h.to(dev), 1.05 * psat)
        p_h.copy_(p)
    flush = torch.empty(L2_FLUSH_BYTES // 8, dtype=torch.float64, device=dev)

    def step_device():
        params.grad = None
        model = PcSaftPure(params)
        nans, vp = model.vapor_pressure(T)
        nans2, rho = model.liquid_density(T, p)
        (vp.sum() + rho.sum()).backward()
        return nans, nans2

    def step_e2e():
        x = params_h.requires_grad_(True)
        x.grad = None
        model = PcSaftPure(x)  # host tensors in: H2D inside
        nans, vp = model.vapor_pressure(T_h)
        nans2, rho = model.liquid_density(T_h, p_h)
        loss = vp.sum() + rho.sum()
        loss.backward()  # gradient lands on the host tensor: D2H inside
        return float(loss), x.grad

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup, collect):
        for _ in range(warmup):
            fn()
        barrier()
        total_ms = 0.0
        if collect:
            _lib.TIMINGS = {}
        for _ in range(steps):
            flush.zero_()  # L2 flush between timed iterations (inputs are smaller than L2)
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            e1.synchronize()
            total_ms += e0.elapsed_time(e1)
        barrier()
        t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    clk_file = tempfile.NamedTemporaryFile(prefix="clocks_", suffix=".csv", delete=False).name
    sampler = clocks_sampler(clk_file, local) if rank == 0 else None
    launches0 = _lib.launch_count()
    ms_dev = timed(step_device, args.steps, args.warmup, True)
    launches = (_lib.launch_count() - launches0) * args.steps // (args.steps + args.warmup)
    timings = _lib.TIMINGS
    _lib.TIMINGS = None
    ms_e2e = timed(step_e2e, args.steps, args.warmup, False)
    if sampler is not None:
        sampler.terminate()
        sampler.wait()

    # failed-state accounting + evaluation counts for the roofline (untimed)
    nans, nans2 = step_device()
    torch.cuda.synchronize()
    failed_frac = nans.double().mean().item()
    evals = torch.zeros(n, dtype=torch.int32, device=dev)
    rho_tmp = torch.empty((n, 2), dtype=torch.float64, device=dev)
    failed_tmp = torch.empty(n, dtype=torch.uint8, device=dev)
    _lib.call("feos_dev_pure_vle_fwd", _lib.ptr(params.detach()), _lib.ptr(T), n, _lib.ptr(rho_tmp), None, None,
              _lib.ptr(failed_tmp), _lib.ptr(evals), torch.cuda.current_stream().cuda_stream)
    mean_evals = evals.double().mean().item()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    kern_ms = {k: sum(a.elapsed_time(b) for a, b in v) / len(v) for k, v in timings.items()}
    kern_steps = {k: [round(a.elapsed_time(b), 3) for a, b in v] for k, v in timings.items()}
    step_ms = ms_dev / args.steps
    dom = "feos_dev_pure_vle_fwd"
    flops_per_launch = n * mean_evals * F_D2
    achieved = flops_per_launch / (kern_ms[dom] * 1e-3) / 1e12
    peak = _lib.measure_fp64_peak()
    spec = 148 * 64 * 2 * 1.965e9 / 1e12
    io_bytes = n * (64 + 8 + 16 + 8 + 1)
    roofline = {
        "bound": "fp64", "kernel": "pure_vle_fwd_kernel(+retry)", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
        "frac": achieved / peak if peak else None, "traffic": None,
        "peak_source": "DFMA-chain microbenchmark in this run (MEASURED_PEAKS.json has no FP64 figure); spec 148 SM x 64 lanes x 2 x 1.965 GHz = %.1f" % spec,
        "flops_per_launch": flops_per_launch, "mean_evals_per_state": mean_evals, "flops_per_eval": F_D2,
        "kernel_ms": kern_ms[dom], "share_of_step": kern_ms[dom] / step_ms,
        "hbm": {"algorithmic_bytes": io_bytes, "achieved_gbs": io_bytes / (kern_ms[dom] * 1e-3) / 1e9, "peak_gbs": _peaks().get("hbm_gbs")},
        "kernel_ms_all": kern_ms, "kernel_ms_per_step": kern_steps,
    }

    value = n * world * args.steps / (ms_dev * 1e-3)
    e2e_value = n * world * args.steps / (ms_e2e * 1e-3)
    cpu = None
    if not args.no_cpu_baseline:
        cpu = cpu_baseline(args.ref_n)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": "C2: PcSaftPure vapor_pressure + liquid_density fwd+bwd, synthetic non-associating (params,T) pairs",
                   "states_per_gpu_per_step": n, "parallelism": f"batch sharded over {world} GPU(s), no data-path collective",
                   "l2": "flushed (256 MiB write) between timed iterations", "failed_fraction": failed_frac},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": n * (64 + 8 + 8) * world, "d2h_bytes_per_step": (n * 64 + 2 * n + 8) * world,
                "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": launches,
        "clocks": parse_clocks(clk_file),
        "roofline": roofline,
        "cpu_baseline": cpu,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def _peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


def cpu_baseline(n):
    from oracle import models as om

    params, T, p = make_workload(n, 1000, "cpu")
    threads = torch.get_num_threads()
    t0 = time.perf_counter()
    x = params.clone().requires_grad_(True)
    model = om.PcSaftPure(x)
    nans, vp = model.vapor_pressure(T)
    psat = torch.zeros(n, dtype=torch.float64)
    psat[~nans] = vp.detach()
    nans2, rho = model.liquid_density(T, torch.maximum(p, 1.05 * psat))
    (vp.sum() + rho.sum()).backward()
    dt = time.perf_counter() - t0
    return {"value": n / dt, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"first {n} states of the C2 workload, one pass, oracle port (torch CPU, {threads} threads), {dt:.1f} s"}


if __name__ == "__main__":
    main()
