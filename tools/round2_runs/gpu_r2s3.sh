#!/bin/bash
# round 2: bev_band hot path through one shared base register, fixed array stride, no per-item divisions
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2s3_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2s3_pytest.log
B="python bench.py --no-e2e --no-cpu-baseline --steps 1500"
ex() { python -c "import json,sys; l=json.loads(sys.stdin.read().strip().splitlines()[-1]); s=l.get('stage_ablation') or {}; k=l['kernels_serialised']; print('$1', round(l['value']), 'single', round((l.get('single_call') or {}).get('value',0)), 'bev_only_us', s.get('bev_only_ms_per_step'), 'bin_ms', k['bev_bin']['ms_per_step'], 'band_ms', k['bev_band']['ms_per_step'])"; }
for rep in 1 2 3; do $B 2>/dev/null | ex "band-trim"; done
$B --config density1r 2>/dev/null | ex "band-trim density1r"
$B --config argoverse 2>/dev/null | ex "band-trim argoverse"
echo -n "single stream ring=32 lanes1: "; SFA_BEV_TILED_RING=32 SFA_BEV_INTERNAL_LANES=1 python tools/bev_run.py 200 3
