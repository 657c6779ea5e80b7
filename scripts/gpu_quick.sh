#!/bin/bash
# Quick A/B on C3: default build + named variants, each with the listed knob settings ("" = defaults). usage: gpu_quick.sh tag "variants" settings...
tag=$1; variants=$2; shift 2
mkdir -p gpurun_out
: > gpurun_out/${tag}_ab.jsonl
timeout 900 python scripts/sweep_tune.py "$@" >> gpurun_out/${tag}_ab.jsonl 2>> gpurun_out/${tag}_ab.err
for v in $variants; do
  echo "== variant '$v'" >> gpurun_out/${tag}_ab.jsonl
  MPT_LIB_VARIANT=$v timeout 600 python scripts/sweep_tune.py "$@" >> gpurun_out/${tag}_ab.jsonl 2>> gpurun_out/${tag}_ab.err
done
cut -c1-330 gpurun_out/${tag}_ab.jsonl; tail -5 gpurun_out/${tag}_ab.err
