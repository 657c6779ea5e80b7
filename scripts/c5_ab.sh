#!/bin/bash
# A/B of two libmpt builds on the instanced config C5 (one GPU). usage: bash scripts/c5_ab.sh variant [variant ...]
for v in "$@"; do echo "== variant '$v'"; MPT_LIB_VARIANT=$v python scripts/bench_configs.py --configs C5 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(round(d['ms_per_frame'],2), round(d['msamples_per_s'],1), round(d['mrays_per_s'],1), d['work'], d['checks'])"; done
