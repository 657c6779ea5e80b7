#!/bin/bash
# Round 2, session D: new node layout (leaf_bits) + cheaper hit-mask assembly + separate node / primitive stacks. GPU tests first.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2d_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2d_pytest.log
tail -15 gpurun_out/r2d_pytest.log
bash scripts/gpu_quick.sh r2d "" "MPT_EXTEND_LEGACY=1" "" "MPT_POOLED_REFILL_THRESHOLD=20" "MPT_POOLED_REFILL_THRESHOLD=28" ""
: > gpurun_out/r2d_c5.jsonl
for leg in 1 0; do
  echo "== C5 legacy=$leg" >> gpurun_out/r2d_c5.jsonl
  MPT_EXTEND_LEGACY=$leg timeout 600 python scripts/bench_configs.py --configs C2,C5 --spp5 16 >> gpurun_out/r2d_c5.jsonl 2>> gpurun_out/r2d_c5.err
done
cut -c1-400 gpurun_out/r2d_c5.jsonl; tail -3 gpurun_out/r2d_c5.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_extend_pooled -s 8 -c 2 -o gpurun_out/r2d_extend \
    python bench.py --steps 1 --warmup 3 --spp 8 --no-cpu-baseline > gpurun_out/r2d_ncu_full.log 2>&1
