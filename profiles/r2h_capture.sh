#!/bin/bash
# ncu evidence of the LAST build of round 2, all four workloads: launch list of the timed device-resident steps of the
# bench command (bench.py brackets them with cudaProfilerStart/Stop when FEOS_BENCH_CUDA_PROFILER=step_device; ncu runs
# with --profile-from-start off) and an --set full capture of the solver kernels, each after the same command has
# exited 0 without ncu.  Exported to CSV on the box (raw page) + executed-flop summaries (flops_from_ncu.py).
set -x
export FEOS_BENCH_CUDA_PROFILER=step_device
cap() {  # cap <workload> <kernel regex> <count>
  B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --extras none --workload $1"
  $B > gpurun_out/r2h_plain_$1.log 2>&1 || return 1
  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2h_launches_bench_$1.csv $B > gpurun_out/r2h_ncu_l_$1.log 2>&1
  ncu --profile-from-start off --set full --clock-control none --import-source on -k "regex:$2" -c $3 -f -o /tmp/r2h_prof_$1 $B > gpurun_out/r2h_ncu_f_$1.log 2>&1
  ncu -i /tmp/r2h_prof_$1.ncu-rep --page raw --csv > gpurun_out/r2h_$1_raw.csv 2>/dev/null
  python profiles/flops_from_ncu.py gpurun_out/r2h_$1_raw.csv > gpurun_out/r2h_$1_summary.txt
  rm -f /tmp/r2h_prof_$1.ncu-rep
}
cap c2 'pure_(vle_fwd|vle_retry|rho_tp_fwd|vle_bwd|rho_tp_bwd)_kernel' 5
cap c3 'pure_(vle_fwd|vle_retry|vle_bwd)_kernel' 3
cap c4 'mix_point_(fwd|retry|bwd_kij)_kernel' 6
cap c5 'gc_point_(fwd|retry|bwd_kab)_kernel|gc_pair_coef' 10
cat gpurun_out/r2h_*_summary.txt
