#!/bin/bash
# Session H (final state of round 1): parity tests, smoke, headline bench, launch list + full captures of extend / shade_primary.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; cut -c1-300 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
T=${TAG:-r01h}
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv \
    python bench.py --steps 1 --warmup 3 --spp 8 --no-cpu-baseline > gpurun_out/${T}_ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_extend -s 16 -c 2 -o gpurun_out/${T}_extend \
    python bench.py --steps 1 --warmup 3 --spp 8 --no-cpu-baseline > gpurun_out/${T}_ncu_full.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_shade_primary -s 4 -c 1 -o gpurun_out/${T}_shade_primary \
    python bench.py --steps 1 --warmup 3 --spp 8 --no-cpu-baseline > gpurun_out/${T}_ncu_full2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_shade_bounce -s 8 -c 1 -o gpurun_out/${T}_shade_bounce \
    python bench.py --steps 1 --warmup 3 --spp 8 --no-cpu-baseline > gpurun_out/${T}_ncu_full3.log 2>&1
ls -la gpurun_out | tail -12
