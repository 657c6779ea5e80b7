#!/bin/bash
# Round 2, session A: GPU tests with the pooled extend kernel, then A/B of legacy vs pooled (+ knobs, + build variants) on C3.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2a_pytest.log
tail -5 gpurun_out/r2a_pytest.log
: > gpurun_out/r2a_ab.jsonl
timeout 900 python scripts/sweep_tune.py "MPT_EXTEND_LEGACY=1" "" "MPT_POOLED_REFILL_THRESHOLD=16" "MPT_POOLED_REFILL_THRESHOLD=20" "MPT_POOLED_REFILL_THRESHOLD=28" \
    "MPT_L2_PERSIST=1" "MPT_L2_PERSIST=1 MPT_EXTEND_LEGACY=1" "MPT_EXTEND_BLOCKS_PER_SM=6" >> gpurun_out/r2a_ab.jsonl 2>> gpurun_out/r2a_ab.err
for v in ready32 ss0 ss8 nostream; do
  echo "== variant '$v'" >> gpurun_out/r2a_ab.jsonl
  MPT_LIB_VARIANT=$v timeout 300 python scripts/sweep_tune.py "" "MPT_POOLED_REFILL_THRESHOLD=20" >> gpurun_out/r2a_ab.jsonl 2>> gpurun_out/r2a_ab.err
done
cat gpurun_out/r2a_ab.jsonl; tail -5 gpurun_out/r2a_ab.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_extend_pooled -s 8 -c 2 -o gpurun_out/r2a_extend \
    python bench.py --steps 1 --warmup 3 --spp 8 --no-cpu-baseline > gpurun_out/r2a_ncu_full.log 2>&1
ls -la gpurun_out | tail -8
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,temperature.gpu --format=csv
