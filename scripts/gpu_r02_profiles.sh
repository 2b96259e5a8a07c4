#!/bin/bash
# r02 evidence: launch list of the multitask epoch and --set full captures of the dominant kernels (profiles/)
mkdir -p gpurun_out
python scripts/prof_step.py multitask_d100 4 > gpurun_out/r02_prof_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r02_launches_multitask_d100_v2.csv \
    python scripts/prof_step.py multitask_d100 4 > gpurun_out/r02_ncu_mt2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:eig_jacobi_oe -s 40 -c 2 -o gpurun_out/r02_prof_eig_oe_mt \
    python scripts/prof_step.py multitask_d100 4 > gpurun_out/r02_ncu_full_eig.log 2>&1
python scripts/prof_step.py single_d100 4 > gpurun_out/r02_prof_plain_single.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:eig_jacobi_oe -s 40 -c 1 -o gpurun_out/r02_prof_eig_oe_single \
    python scripts/prof_step.py single_d100 4 > gpurun_out/r02_ncu_full_eig_single.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02_launches_single_d100_v2.csv \
    python scripts/prof_step.py single_d100 4 > gpurun_out/r02_ncu_single2.log 2>&1
