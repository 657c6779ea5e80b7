#!/bin/bash
# Round 2, session P: ncu --set full of the two-level extend kernel on C5 (instanced rocks), source page included, + the same for C3 (final
# single-level kernel of this session).
mkdir -p gpurun_out
SWEEP_SCENE=c5 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_extend_pooled -s 9 -c 2 -o gpurun_out/r2p_c5_extend \
    python scripts/sweep_tune.py "" > gpurun_out/r2p_c5_ncu.log 2>&1
tail -3 gpurun_out/r2p_c5_ncu.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_extend_pooled -s 8 -c 2 -o gpurun_out/r2p_c3_extend \
    python bench.py --steps 1 --warmup 3 --spp 8 --no-cpu-baseline > gpurun_out/r2p_c3_ncu.log 2>&1
tail -3 gpurun_out/r2p_c3_ncu.log
ls -la gpurun_out | tail -4
