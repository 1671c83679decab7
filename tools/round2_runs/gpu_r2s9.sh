#!/bin/bash
# round 2: engines x internal lanes with programmatic launches on both boundaries
mkdir -p gpurun_out
B="timeout 300 python bench.py --no-e2e --no-cpu-baseline --no-stage-ablation --steps 1500"
ex() { python -c "import json,sys; l=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', round(l['value']), 'single', round((l.get('single_call') or {}).get('value',0)))"; }
for p in 1 2 3 4; do
$B --pipelines $p 2>/dev/null | ex "uniform engines=$p"
done
for p in 2 3 4; do
$B --pipelines $p --config density1r 2>/dev/null | ex "density1r engines=$p"
done
for p in 2 3; do
$B --pipelines $p --config argoverse 2>/dev/null | ex "argoverse engines=$p"
done
for l in 1 2 3; do echo -n "one stream, internal lanes=$l: "; SFA_BEV_INTERNAL_LANES=$l timeout 120 python tools/bev_run.py 200 3; done
