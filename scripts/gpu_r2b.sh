#!/bin/bash
# Round 2, session B: pooled extend with 48-byte records and the stack in local memory; knobs (refill threshold, primitive-section
# period), pool sizes, C5 (two-level) legacy vs pooled.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2b_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2b_pytest.log
tail -3 gpurun_out/r2b_pytest.log
: > gpurun_out/r2b_ab.jsonl
timeout 900 python scripts/sweep_tune.py "MPT_EXTEND_LEGACY=1" "" "MPT_POOLED_TRI_MASK=1" "MPT_POOLED_TRI_MASK=3" "MPT_POOLED_REFILL_THRESHOLD=20" "MPT_POOLED_REFILL_THRESHOLD=28" \
    "MPT_POOLED_REFILL_THRESHOLD=28 MPT_POOLED_TRI_MASK=1" "" >> gpurun_out/r2b_ab.jsonl 2>> gpurun_out/r2b_ab.err
for v in ready8 ready24 stream; do
  echo "== variant '$v'" >> gpurun_out/r2b_ab.jsonl
  MPT_LIB_VARIANT=$v timeout 300 python scripts/sweep_tune.py "" >> gpurun_out/r2b_ab.jsonl 2>> gpurun_out/r2b_ab.err
done
cat gpurun_out/r2b_ab.jsonl; tail -5 gpurun_out/r2b_ab.err
: > gpurun_out/r2b_c5.jsonl
for leg in 1 0; do
  echo "== C5 legacy=$leg" >> gpurun_out/r2b_c5.jsonl
  MPT_EXTEND_LEGACY=$leg timeout 600 python scripts/bench_configs.py --configs C5 --spp5 16 >> gpurun_out/r2b_c5.jsonl 2>> gpurun_out/r2b_c5.err
done
cat gpurun_out/r2b_c5.jsonl; tail -3 gpurun_out/r2b_c5.err
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,temperature.gpu --format=csv
