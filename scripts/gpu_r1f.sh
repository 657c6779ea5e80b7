#!/bin/bash
# Session F: GPU parity tests, ray-ordering experiment, run-time knob sweep of k_extend (tri vote / prefetch / refill).
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpu.txt; nproc >> gpurun_out/gpu.txt
timeout 1200 python -m pytest tests -m gpu -x -q --durations=5 > gpurun_out/pytest_gpu.log 2>&1; tail -8 gpurun_out/pytest_gpu.log
timeout 300 python scripts/exp_ray_sorting.py > gpurun_out/ray_sorting.jsonl 2> gpurun_out/ray_sorting.err; cat gpurun_out/ray_sorting.jsonl; tail -3 gpurun_out/ray_sorting.err
timeout 900 python scripts/sweep_tune.py "" "MPT_TRI_VOTE=4" "MPT_TRI_VOTE=8" "MPT_TRI_VOTE=12" "MPT_TRI_VOTE=16" "MPT_TRI_VOTE=24" \
  "MPT_PREFETCH=1" "MPT_PREFETCH=2" "MPT_PREFETCH=3" "MPT_TRI_VOTE=8 MPT_PREFETCH=3" "MPT_TRI_VOTE=8 MPT_PREFETCH=2" \
  "MPT_TRI_VOTE=8 MPT_REFILL_THRESHOLD=8" "MPT_TRI_VOTE=8 MPT_REFILL_THRESHOLD=12" "MPT_TRI_VOTE=8 MPT_REFILL_THRESHOLD=20" "MPT_TRI_VOTE=8 MPT_REFILL_THRESHOLD=24" \
  "MPT_REFILL_THRESHOLD=24" "" > gpurun_out/sweep_tune.jsonl 2> gpurun_out/sweep_tune.err
cat gpurun_out/sweep_tune.jsonl; tail -3 gpurun_out/sweep_tune.err
