"""Executed FP64 flops per launch from an `ncu --set full ... --page raw --csv` export:
(2 DFMA + DMUL + DADD thread instructions per elapsed cycle, summed over the SMSPs) x elapsed cycles.

    python profiles/flops_from_ncu.py profiles/r2f_mix_raw.csv [states evals_per_state]
"""
import csv
import sys


def kernels(path):
    rows = list(csv.reader(open(path)))
    hdr = rows[0]
    for r in rows[2:]:
        d = dict(zip(hdr, r))

        def g(k):
            try:
                return float(d.get(k, "") or 0.0)
            except ValueError:
                return 0.0
        per_cycle = (2.0 * g("smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed")
                     + g("smsp__sass_thread_inst_executed_op_dmul_pred_on.sum.per_cycle_elapsed")
                     + g("smsp__sass_thread_inst_executed_op_dadd_pred_on.sum.per_cycle_elapsed"))
        cycles = g("smsp__cycles_elapsed.max") or g("sm__cycles_elapsed.max")
        unit = rows[1][hdr.index("gpu__time_duration.sum")]
        t = g("gpu__time_duration.sum") * {"us": 1e-6, "ms": 1e-3, "ns": 1e-9, "s": 1.0}.get(unit, 1e-6)
        yield {
            "kernel": d["Kernel Name"].split("(")[0], "seconds": t, "cycles": cycles, "flops": per_cycle * cycles,
            "flops_per_cycle": per_cycle, "fp64_pipe_pct": g("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
            "lanes": g("smsp__thread_inst_executed_per_inst_executed.ratio"), "regs": g("launch__registers_per_thread"),
            "icc_hit_pct": g("sm__icc_request_hit_rate.pct"),
            "dram_mb": sum(g(k) * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(rows[1][hdr.index(k)], 1.0)
                           for k in ("dram__bytes_read.sum", "dram__bytes_write.sum") if k in hdr),
            "dram_unit": "MB",
        }


if __name__ == "__main__":
    n = float(sys.argv[2]) if len(sys.argv) > 2 else None
    ev = float(sys.argv[3]) if len(sys.argv) > 3 else None
    for k in kernels(sys.argv[1]):
        line = ("%-34s %8.1f us  %5.0f regs  fp64 pipe %5.1f %%  lanes %4.1f  icc hit %5.1f %%  flops %.4e  (%.0f / cycle = %.3f of 18944)  "
                "%.2f TFLOP/s  dram %.1f %s" % (k["kernel"][:34], k["seconds"] * 1e6, k["regs"], k["fp64_pipe_pct"], k["lanes"], k["icc_hit_pct"],
                                               k["flops"], k["flops_per_cycle"], k["flops_per_cycle"] / 18944.0,
                                               k["flops"] / k["seconds"] / 1e12 if k["seconds"] else 0.0, k["dram_mb"], k["dram_unit"]))
        if n and ev:
            line += "  -> %.0f flops per state-evaluation" % (k["flops"] / (n * ev))
        print(line)
