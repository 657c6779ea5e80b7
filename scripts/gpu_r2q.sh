#!/bin/bash
# Round 2, session Q: shared-reciprocal shear quotients, arithmetic make_perm, 1-ulp inverse direction at instance entry, flat pop section for
# two-level scenes (default) vs the loop form (variant poploop, same other changes). Full GPU tests first.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/r2q_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2q_pytest.log
tail -6 gpurun_out/r2q_pytest.log
SWEEP_SCENE=c5 bash scripts/gpu_quick.sh r2q_c5 "poploop" "" ""
bash scripts/gpu_quick.sh r2q "poploop" "" ""
