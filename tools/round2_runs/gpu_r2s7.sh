#!/bin/bash
# round 2: 512 band CTAs + half-frame band rotation between a CTA's visits + programmatic dependent launch (defaults now)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2s7_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2s7_pytest.log
B="timeout 300 python bench.py --no-e2e --no-cpu-baseline --steps 1500"
ex() { python -c "import json,sys; l=json.loads(sys.stdin.read().strip().splitlines()[-1]); s=l.get('stage_ablation') or {}; k=l['kernels_serialised']; print('$1', round(l['value']), 'single', round((l.get('single_call') or {}).get('value',0)), 'bev_only_us', s.get('bev_only_ms_per_step'), 'band_ms', k['bev_band']['ms_per_step'])"; }
for rep in 1 2; do $B 2>/dev/null | ex "default"; done
SFA_BAND_FILL_SLOTS=1 $B 2>/dev/null | ex "fill"
SFA_BEV_PDL=0 $B 2>/dev/null | ex "nopdl"
$B --config density1r 2>/dev/null | ex "default density1r"
SFA_BAND_FILL_SLOTS=1 $B --config density1r 2>/dev/null | ex "fill density1r"
$B --config argoverse 2>/dev/null | ex "default argoverse"
SFA_BAND_FILL_SLOTS=1 $B --config argoverse 2>/dev/null | ex "fill argoverse"
timeout 300 python tools/bev_distributions.py 2>&1 | tail -5
