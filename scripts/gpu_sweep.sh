#!/bin/bash
# Quick tuning sweep on one B200: parity tests first, then bench.py (no CPU baseline) for each "VAR=val ..." line of $1.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log
: > gpurun_out/sweep.jsonl
while IFS= read -r line; do
  [ -z "$line" ] && continue
  echo "== $line" >> gpurun_out/sweep.jsonl
  env $line timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline >> gpurun_out/sweep.jsonl 2>> gpurun_out/sweep.err
done < "$1"
python - <<'PY'
import json
for l in open('gpurun_out/sweep.jsonl'):
    if l.startswith('=='): print(l.strip(), end='  ')
    else:
        try:
            d=json.loads(l); r=d['roofline']; print(f"{d['value']:.0f} Ms/s  extend {r['avg_launch_ms']:.3f} ms  share {r['extend_share_of_step']:.2f} shade {r['shade_share_of_step']:.2f} frac {r['frac']:.3f} clk {d['clocks']['sm_mhz']} {d['clocks']['reasons']}")
        except Exception as e: print('ERR', l[:200])
PY
