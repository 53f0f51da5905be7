#!/bin/bash
# ncu evidence for the state-order change (run under gpurun): launch list of the C2 bench command + --set full
# captures of the forward kernels in state order and of the GC backward kernel.
set -x
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$B --workload c2 > gpurun_out/plain_c2.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1d_c2.csv $B --workload c2 > gpurun_out/ncu_l_c2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:pure_vle_fwd_kernel -s 2 -c 1 -f -o gpurun_out/prof_r1d_pure_vle_fwd $B --workload c2 > gpurun_out/ncu_f_c2.log 2>&1
$B --workload c4 > gpurun_out/plain_c4.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:mix_point_fwd_kernel -s 2 -c 2 -f -o gpurun_out/prof_r1d_mix_point_fwd $B --workload c4 > gpurun_out/ncu_f_c4.log 2>&1
$B --workload c5 > gpurun_out/plain_c5.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k 'regex:gc_point_(fwd|bwd)_kernel' -s 4 -c 3 -f -o gpurun_out/prof_r1d_gc_point $B --workload c5 > gpurun_out/ncu_f_c5.log 2>&1
ls -la gpurun_out/*r1d*
