#!/bin/bash
# Round-1 GPU session: tests, smoke, bench, ncu launch list + one full capture of k_extend / k_shade_bounce.
set -x
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpu.txt; nproc >> gpurun_out/gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -15 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -3 gpurun_out/smoke.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -c 3000 gpurun_out/bench.json
tail -5 gpurun_out/bench.err
# launch list: every launch of a short run (cold-cache, serialised: compare shares)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 3 --spp 8 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
# full capture of the extend kernel (3 launches after warm-up)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_extend -s 8 -c 3 -o gpurun_out/prof_extend \
    python bench.py --steps 1 --warmup 3 --spp 8 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_shade_bounce -s 8 -c 2 -o gpurun_out/prof_shade \
    python bench.py --steps 1 --warmup 3 --spp 8 --no-cpu-baseline > gpurun_out/ncu_full2.log 2>&1
ls -la gpurun_out
