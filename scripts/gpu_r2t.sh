#!/bin/bash
# Round 2, session T: L1 policy of the triangle-record / node loads (variants trina, trief, nodeel, both) and the refill threshold after the
# session's instruction-count changes, on C3; the best two also on C5.
mkdir -p gpurun_out
bash scripts/gpu_quick.sh r2t "trina trief nodeel both" ""
timeout 600 python scripts/sweep_tune.py "MPT_POOLED_REFILL_THRESHOLD=10" "MPT_POOLED_REFILL_THRESHOLD=18" "MPT_POOLED_REFILL_THRESHOLD=22" > gpurun_out/r2t_refill.jsonl 2> gpurun_out/r2t_refill.err; cut -c1-200 gpurun_out/r2t_refill.jsonl
SWEEP_SCENE=c5 bash scripts/gpu_quick.sh r2t_c5 "trina both" ""
