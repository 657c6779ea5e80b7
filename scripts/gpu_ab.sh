#!/bin/bash
# A/B of two builds of libmpt (MPT_LIB_VARIANT) on C3: same box, same session, alternating.
mkdir -p gpurun_out
: > gpurun_out/ab.jsonl
for v in "$@"; do
  echo "== variant '$v'" >> gpurun_out/ab.jsonl
  MPT_LIB_VARIANT=$v timeout 600 python scripts/sweep_tune.py "" >> gpurun_out/ab.jsonl 2>> gpurun_out/ab.err
done
cat gpurun_out/ab.jsonl; tail -3 gpurun_out/ab.err
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,temperature.gpu --format=csv
