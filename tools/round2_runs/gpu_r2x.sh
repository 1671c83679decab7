#!/bin/bash
# round 2: fewer, larger bands (less per-item and per-run overhead) in the timed schedule
mkdir -p gpurun_out
B="python bench.py --no-e2e --no-cpu-baseline --steps 1500"
ex() { python -c "import json,sys; l=json.loads(sys.stdin.read().strip().splitlines()[-1]); s=l.get('stage_ablation') or {}; print('$1', round(l['value']), 'single', round((l.get('single_call') or {}).get('value',0)), 'bev_only_us', s.get('bev_only_ms_per_step'))"; }
for nb in 128 112 108 96 80 64; do
SFA_BEV_BANDS=$nb python -m pytest tests/test_bev_gpu.py -m gpu -x -q -k "batch64 or fixture" 2>&1 | tail -1
SFA_BEV_BANDS=$nb $B 2>/dev/null | ex "bands$nb"
SFA_BEV_BANDS=$nb $B --config density1r 2>/dev/null | ex "bands$nb density1r"
done
