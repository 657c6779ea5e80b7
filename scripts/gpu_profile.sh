#!/bin/bash
# ncu session on one B200: launch list of a short bench run + full captures of the extend and shade kernels.
# usage: bash scripts/gpu_profile.sh <tag>   -> gpurun_out/<tag>_*.{csv,ncu-rep}
T=${1:-prof}
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv \
    python bench.py --steps 1 --warmup 3 --spp 8 --no-cpu-baseline > gpurun_out/${T}_ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_extend -s 8 -c 2 -o gpurun_out/${T}_extend \
    python bench.py --steps 1 --warmup 3 --spp 8 --no-cpu-baseline > gpurun_out/${T}_ncu_full.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_shade_bounce -s 8 -c 1 -o gpurun_out/${T}_shade \
    python bench.py --steps 1 --warmup 3 --spp 8 --no-cpu-baseline > gpurun_out/${T}_ncu_full2.log 2>&1
ls -la gpurun_out
