#!/bin/bash
# C3 (10 M associating states): the workload generator (a bisection on the critical temperature: ~60 solves with ever
# more near-critical states) launches the same kernels, so ncu is restricted to the timed device-resident steps
# (bench.py calls cudaProfilerStart/Stop around them when FEOS_BENCH_CUDA_PROFILER=step_device).
set -x
export FEOS_BENCH_CUDA_PROFILER=step_device
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --extras none --workload c3"
$B > gpurun_out/r2f_plain_c3.log 2>&1 || exit 1
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2f_launches_bench_c3.csv $B > gpurun_out/r2f_ncu_l_c3.log 2>&1
ncu --profile-from-start off --set full --clock-control none --import-source on -k "regex:pure_(vle_fwd|vle_retry|vle_bwd)_kernel" -c 3 -f -o /tmp/r2f_prof_c3 $B > gpurun_out/r2f_ncu_f_c3.log 2>&1
ncu -i /tmp/r2f_prof_c3.ncu-rep --page raw --csv > gpurun_out/r2f_pure_c3_raw.csv 2>/dev/null
python profiles/flops_from_ncu.py gpurun_out/r2f_pure_c3_raw.csv 1e7 12.145 > gpurun_out/r2f_pure_c3_summary.txt
cat gpurun_out/r2f_pure_c3_summary.txt
