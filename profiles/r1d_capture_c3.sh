#!/bin/bash
# ncu --set full of the pure forward / backward kernels on the associating C3 workload (run under gpurun)
set -x
B="python bench.py --workload c3 --steps 2 --warmup 3 --no-cpu-baseline"
$B > gpurun_out/plain_c3.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k 'regex:pure_vle_(fwd|bwd)_kernel' -s 26 -c 2 -f -o gpurun_out/prof_r1d_pure_c3 $B > gpurun_out/ncu_f_c3.log 2>&1
ls -la gpurun_out/prof_r1d_pure_c3.ncu-rep
