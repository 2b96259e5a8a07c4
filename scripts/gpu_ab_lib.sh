#!/bin/bash
# A/B of two builds of the library on one box: uglad_b200/lib/libuglad_b200_base.so (copied before a change) vs the current one
for rep in 1 2; do
  for lib in base new; do
    if [ $lib = base ]; then export UGLAD_B200_LIB=$PWD/uglad_b200/lib/libuglad_b200_base.so; else unset UGLAD_B200_LIB; fi
    timeout 200 python bench.py --steps 20 --warmup 5 --no-extra 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('$lib multitask', round(d['value']), round(d['ms_per_step'],3), 'eig avg launch ms', round(r['avg_launch_ms'],4), 'share', round(r['kernel_share_of_step'],3))"
    timeout 200 python bench.py --steps 20 --warmup 5 --no-extra --workload single_d100 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('$lib single_d100', round(d['value']), round(d['ms_per_step'],3), 'eig avg launch ms', round(r['avg_launch_ms'],4))"
  done
done
