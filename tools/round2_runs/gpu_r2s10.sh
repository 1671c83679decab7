#!/bin/bash
# round 2: bev_bin lets the dependent bev_band move in as soon as every bin CTA has started (SFA_BIN_EARLY_DEPENDENTS=1) vs after staging
mkdir -p gpurun_out
SFA_BIN_EARLY_DEPENDENTS=1 timeout 300 python -m pytest tests/test_bev_gpu.py -m gpu -x -q -k "batch64 or fixture or graph or lanes or random_geometries or front_back or overflow" > gpurun_out/r2s10_pytest.log 2>&1; echo "pytest(early) rc=$?"; tail -1 gpurun_out/r2s10_pytest.log
B="timeout 200 python bench.py --no-e2e --no-cpu-baseline --no-stage-ablation --steps 1500"
ex() { python -c "import json,sys; l=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', round(l['value']), 'single', round((l.get('single_call') or {}).get('value',0)))"; }
for rep in 1 2; do
$B 2>/dev/null | ex "late"
SFA_BIN_EARLY_DEPENDENTS=1 $B 2>/dev/null | ex "early"
done
SFA_BIN_EARLY_DEPENDENTS=1 $B --config density1r 2>/dev/null | ex "early density1r"
for e in 0 1; do echo -n "one stream, lanes 1, early=$e: "; SFA_BIN_EARLY_DEPENDENTS=$e SFA_BEV_INTERNAL_LANES=1 timeout 100 python tools/bev_run.py 200 3; done
