#!/bin/bash
# Round 2, session N: branch-free component selection in the watertight test + one relative guard in the node test. GPU tests, then A/B
# against the build with the previous forms (variant "ternary") on C3 and C5.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/r2n_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2n_pytest.log
tail -6 gpurun_out/r2n_pytest.log
bash scripts/gpu_quick.sh r2n "ternary" "" ""
SWEEP_SCENE=c5 bash scripts/gpu_quick.sh r2n_c5 "ternary" ""
