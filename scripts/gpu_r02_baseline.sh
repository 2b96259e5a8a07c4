#!/bin/bash
# r02 baseline: where the time goes per workload before this round's changes
set -x
mkdir -p gpurun_out
python scripts/gpu_eig_phases.py 256 100 > gpurun_out/r02_phases_b256.log 2>&1
python scripts/gpu_eig_phases.py 1 100 > gpurun_out/r02_phases_b1.log 2>&1
python scripts/gpu_eig_phases.py 32 100 > gpurun_out/r02_phases_b32.log 2>&1
python scripts/gpu_time.py > gpurun_out/r02_gpu_time.log 2>&1
python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench0.json 2> gpurun_out/r02_bench0.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r02_launches_single_d100.csv python scripts/prof_step.py single_d100 4 > gpurun_out/r02_ncu_single.log 2>&1
