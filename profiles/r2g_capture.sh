#!/bin/bash
# Round 2, last build: GPU tests, the default bench command (headline C2 + C3 / C4 / C5 legs + CPU baseline), the
# reference arm, and a re-capture of the C5 kernels (changed after r2f_capture.sh: DipolePoly, per-pair table).
set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r2g_gpu_tests.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" >> gpurun_out/r2g_gpu_tests.log 2>&1
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2g_bench_reference.json 2> gpurun_out/r2g_bench_reference.err
python bench.py > gpurun_out/r2g_bench_default.json 2> gpurun_out/r2g_bench_default.err
export FEOS_BENCH_CUDA_PROFILER=step_device
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --extras none --workload c5"
$B > gpurun_out/r2g_plain_c5.log 2>&1 || exit 1
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2g_launches_bench_c5.csv $B > gpurun_out/r2g_ncu_l_c5.log 2>&1
ncu --profile-from-start off --set full --clock-control none --import-source on -k "regex:gc_point_(fwd|bwd_kab)_kernel|gc_pair_coef" -c 9 -f -o /tmp/r2g_prof_gc $B > gpurun_out/r2g_ncu_f_c5.log 2>&1
ncu -i /tmp/r2g_prof_gc.ncu-rep --page raw --csv > gpurun_out/r2g_gc_raw.csv 2>/dev/null
python profiles/flops_from_ncu.py gpurun_out/r2g_gc_raw.csv > gpurun_out/r2g_gc_summary.txt
cat gpurun_out/r2g_gc_summary.txt
tail -3 gpurun_out/r2g_gpu_tests.log
