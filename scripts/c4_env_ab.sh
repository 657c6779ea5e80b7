#!/bin/bash
# A/B of environment knobs on C5. usage: bash scripts/c5_env_ab.sh "VAR=1" "" ...
for v in "$@"; do echo "== env '$v'"; env $v python scripts/bench_configs.py --configs C4 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(round(d['ms_per_frame'],2), round(d['msamples_per_s'],1), round(d['mrays_per_s'],1), d['work'], d['checks'], round(d['upload_s'],2))"; done
