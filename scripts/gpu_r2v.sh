#!/bin/bash
# Round 2, session V: A/B of libmpt variants on C3 and C5 (one GPU, same box). usage: bash scripts/gpu_r2v.sh "c3 variants" "c5 variants"
mkdir -p gpurun_out
: > gpurun_out/r2v_c3.jsonl; : > gpurun_out/r2v_c5.jsonl
for v in $1; do
  [ "$v" = base ] && v=""
  echo "== variant '$v'" >> gpurun_out/r2v_c3.jsonl
  MPT_LIB_VARIANT=$v timeout 600 python scripts/sweep_tune.py "" >> gpurun_out/r2v_c3.jsonl 2>> gpurun_out/r2v.err
done
for v in $2; do
  [ "$v" = base ] && v=""
  echo "== variant '$v'" >> gpurun_out/r2v_c5.jsonl
  SWEEP_SCENE=c5 MPT_LIB_VARIANT=$v timeout 600 python scripts/sweep_tune.py "" >> gpurun_out/r2v_c5.jsonl 2>> gpurun_out/r2v.err
done
cut -c1-330 gpurun_out/r2v_c3.jsonl; cut -c1-330 gpurun_out/r2v_c5.jsonl; tail -3 gpurun_out/r2v.err
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,temperature.gpu --format=csv
