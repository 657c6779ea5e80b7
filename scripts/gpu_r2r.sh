#!/bin/bash
# Round 2, session R (measurement of the session's final build, one B200): GPU tests, smoke, bench line, READY=12 A/B, launch list + ncu
# --set full capture of the timed step's extend launches (profiles/extend_traffic.json), the other BASELINE configs, reference arm.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2r_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2r_pytest.log
tail -5 gpurun_out/r2r_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2r_bench.json 2> gpurun_out/r2r_bench.err; cut -c1-400 gpurun_out/r2r_bench.json; tail -3 gpurun_out/r2r_bench.err
bash scripts/gpu_quick.sh r2r "ready12" "" ""
bash scripts/gpu_traffic.sh r2r > gpurun_out/r2r_traffic.log 2>&1; tail -3 gpurun_out/r2r_traffic.log
timeout 900 python scripts/bench_configs.py --configs C1,C2,C5 --spp5 16 > gpurun_out/r2r_configs.jsonl 2> gpurun_out/r2r_configs.err; cut -c1-500 gpurun_out/r2r_configs.jsonl; tail -3 gpurun_out/r2r_configs.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2r_bench_reference.json 2> gpurun_out/r2r_bench_reference.err; cut -c1-600 gpurun_out/r2r_bench_reference.json
