#!/bin/bash
# Round 2, last build: GPU tests, smoke, reference arm, the default bench command.
python -m pytest tests -m gpu -x -q > gpurun_out/r2h_gpu_tests.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" >> gpurun_out/r2h_gpu_tests.log 2>&1
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2h_bench_reference.json 2> gpurun_out/r2h_bench_reference.err
python bench.py > gpurun_out/r2h_bench_default.json 2> gpurun_out/r2h_bench_default.err
tail -3 gpurun_out/r2h_gpu_tests.log
