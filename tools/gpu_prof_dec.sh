#!/bin/bash
# ncu --set full of decode kernels at a reduced bench size. usage: tools/gpu_prof_dec.sh <tag> <kernel-regex> [launch-skip] [count]
tag=$1; rx=$2; skip=${3:-1}; cnt=${4:-1}
ncu --set full --clock-control none --import-source on -k regex:"$rx" --launch-skip $skip -c $cnt -o gpurun_out/${tag} -f \
    python bench.py --steps 1 --warmup 1 --tracks-per-gpu 32 --no-cpu-baseline --no-e2e > gpurun_out/${tag}.log 2>&1
tail -3 gpurun_out/${tag}.log | cut -c1-300
