#!/bin/bash
# Round 2, session Y (measurement of the round's final build, one B200): GPU tests, smoke, bench line, end-to-end probe (host stalls), launch
# list + ncu --set full capture of the timed step's extend launches, the other BASELINE configs, reference arm.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2y_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2y_pytest.log
tail -4 gpurun_out/r2y_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r2y_bench.json 2> gpurun_out/r2y_bench.err; cut -c1-400 gpurun_out/r2y_bench.json; tail -3 gpurun_out/r2y_bench.err
timeout 300 python scripts/e2e_probe.py 40 > gpurun_out/r2y_e2e_probe.json 2> gpurun_out/r2y_e2e_probe.err
PROBE_GC_OFF=1 timeout 300 python scripts/e2e_probe.py 40 > gpurun_out/r2y_e2e_probe_gcoff.json 2>> gpurun_out/r2y_e2e_probe.err
bash scripts/gpu_traffic.sh r2y > gpurun_out/r2y_traffic.log 2>&1; tail -3 gpurun_out/r2y_traffic.log
timeout 1500 python scripts/bench_configs.py --configs C1,C2,C4,C5 --spp5 16 > gpurun_out/r2y_configs.jsonl 2> gpurun_out/r2y_configs.err; cut -c1-500 gpurun_out/r2y_configs.jsonl; tail -3 gpurun_out/r2y_configs.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2y_bench_reference.json 2> gpurun_out/r2y_bench_reference.err; cut -c1-600 gpurun_out/r2y_bench_reference.json
