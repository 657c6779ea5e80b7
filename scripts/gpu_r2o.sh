#!/bin/bash
# Round 2, session O: thin-lens test, flat pop section (default) vs the loop form (variant poploop), READY pool of 12 / 10 records (built on the loop form).
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_features.py tests/test_gpu_parity.py tests/test_abi.py -m gpu -q -x > gpurun_out/r2o_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2o_pytest.log
tail -6 gpurun_out/r2o_pytest.log
bash scripts/gpu_quick.sh r2o "poploop ready12 ready10" "" ""
SWEEP_SCENE=c5 bash scripts/gpu_quick.sh r2o_c5 "poploop" ""
