#!/bin/bash
# Round 2, session S: where the pipelined host-buffer path spends its time (e2e probe); full-size parity of the session's final build.
mkdir -p gpurun_out
timeout 600 python scripts/e2e_probe.py 20 > gpurun_out/r2s_e2e_probe.json 2> gpurun_out/r2s_e2e_probe.err; cut -c1-1500 gpurun_out/r2s_e2e_probe.json; tail -3 gpurun_out/r2s_e2e_probe.err
timeout 1500 python scripts/parity_fullsize.py --configs C2,C5,C3,C4 --out gpurun_out/r2s_parity_fullsize.json > gpurun_out/r2s_parity.log 2>&1; tail -5 gpurun_out/r2s_parity.log
