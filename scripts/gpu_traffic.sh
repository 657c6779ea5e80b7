#!/bin/bash
# ncu --set full capture of the 8 extend launches of the timed step of the default bench command (for profiles/extend_traffic.json),
# plus the launch list of the same command (gpu__time_duration only: the kernels' SHARE of a step, not their absolute times).
# Launches before the timed step with --steps 1 --warmup 3 --no-cpu-baseline: 8 (counting frame) + 6 * 8 (pre-warm) + 3 * 8 (warm-up) = 80.
T=${1:-r02}
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_extend -s 80 -c 8 -o gpurun_out/${T}_extend_step \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_ncu_step.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_ncu_launches.log 2>&1
ls -la gpurun_out | tail -5
