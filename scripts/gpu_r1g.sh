#!/bin/bash
# Session G (end of round 1): parity tests, smoke, headline bench, reference arm, ncu launch list + full captures, the other BASELINE configs,
# 4096-spp radiance parity, pool-size check. Everything lands in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpu.txt; nproc >> gpurun_out/gpu.txt
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; cut -c1-700 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2>> gpurun_out/bench.err; cut -c1-400 gpurun_out/bench_reference.json
timeout 600 python scripts/parity_4096spp.py > gpurun_out/parity.log 2>&1; tail -2 gpurun_out/parity.log
timeout 600 python scripts/sweep_tune.py "" "MPT_POOL_MSLOTS=128" "MPT_POOL_MSLOTS=32" > gpurun_out/sweep_pool.jsonl 2> gpurun_out/sweep_pool.err; cut -c1-200 gpurun_out/sweep_pool.jsonl
timeout 900 python scripts/bench_configs.py --configs C1,C2,C5 > gpurun_out/configs.jsonl 2> gpurun_out/configs.err; cut -c1-300 gpurun_out/configs.jsonl; tail -2 gpurun_out/configs.err
T=r01g
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv \
    python bench.py --steps 1 --warmup 3 --spp 8 --no-cpu-baseline > gpurun_out/${T}_ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_extend -s 8 -c 3 -o gpurun_out/${T}_extend \
    python bench.py --steps 1 --warmup 3 --spp 8 --no-cpu-baseline > gpurun_out/${T}_ncu_full.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_shade_bounce -s 8 -c 1 -o gpurun_out/${T}_shade \
    python bench.py --steps 1 --warmup 3 --spp 8 --no-cpu-baseline > gpurun_out/${T}_ncu_full2.log 2>&1
ls -la gpurun_out | tail -30
