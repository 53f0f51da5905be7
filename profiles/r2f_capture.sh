#!/bin/bash
# ncu evidence, round 2 final kernels (run under gpurun): launch list of the bench command per workload and --set full
# captures of the dominant kernels, each after the same command has exited 0 without ncu.  The .ncu-rep files are
# exported to CSV on the box (raw metrics per launch; per-instruction SASS view gzipped) and removed (size limit).
set -x
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --extras none"
cap() {  # cap <workload> <name> <kernel regex> <skip> <count>
  $B --workload $1 > gpurun_out/r2f_plain_$1.log 2>&1 || return 1
  ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r2f_launches_bench_$1.csv $B --workload $1 > gpurun_out/r2f_ncu_l_$1.log 2>&1
  ncu --set full --clock-control none --import-source on -k "regex:$3" -s $4 -c $5 -f -o /tmp/r2f_prof_$2 $B --workload $1 > gpurun_out/r2f_ncu_f_$1.log 2>&1
  ncu -i /tmp/r2f_prof_$2.ncu-rep --page raw --csv > gpurun_out/r2f_$2_raw.csv 2>/dev/null
  ncu -i /tmp/r2f_prof_$2.ncu-rep --page source --csv --print-source sass > gpurun_out/r2f_$2_sass.csv 2>/dev/null
  rm -f /tmp/r2f_prof_$2.ncu-rep
}
cap c2 pure 'pure_(vle_fwd|vle_retry|rho_tp_fwd|vle_bwd|rho_tp_bwd)_kernel' 15 5
cap c3 pure_c3 'pure_(vle_fwd|vle_retry|vle_bwd)_kernel' 9 3
cap c4 mix 'mix_point_(fwd|bwd_kij)_kernel' 24 8
cap c5 gc 'gc_point_(fwd|bwd_kab)_kernel' 18 6
gzip -f gpurun_out/r2f_*_sass.csv
for f in gpurun_out/r2f_*_raw.csv; do python profiles/flops_from_ncu.py $f > ${f%_raw.csv}_summary.txt; done
ls -la gpurun_out/r2f_*
