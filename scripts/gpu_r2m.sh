#!/bin/bash
# Round 2, session M: full GPU test run (no -x) + smoke + bench line on the current build.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --durations=5 > gpurun_out/r2m_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2m_pytest.log
tail -25 gpurun_out/r2m_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2m_bench.json 2> gpurun_out/r2m_bench.err; cat gpurun_out/r2m_bench.json; tail -3 gpurun_out/r2m_bench.err
