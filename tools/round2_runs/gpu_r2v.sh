#!/bin/bash
# round 2: ping-pong band kernel (SFA_BEV_BAND_PP=1) against the default
mkdir -p gpurun_out
SFA_BEV_BAND_PP=1 timeout 600 python -m pytest tests/test_bev_gpu.py tests/test_augment_gpu.py -m gpu -x -q > gpurun_out/r2v_pytest.log 2>&1; echo "pytest(pp) rc=$?"; tail -3 gpurun_out/r2v_pytest.log
B="python bench.py --no-e2e --no-cpu-baseline --steps 1500"
ex() { python -c "import json,sys; l=json.loads(sys.stdin.read().strip().splitlines()[-1]); s=l.get('stage_ablation') or {}; k=l['kernels_serialised']; print('$1', round(l['value']), 'single', round((l.get('single_call') or {}).get('value',0)), 'bev_only_us', s.get('bev_only_ms_per_step'), 'band_ms', k['bev_band']['ms_per_step'])"; }
for rep in 1 2; do
$B 2>/dev/null | ex "default"
SFA_BEV_BAND_PP=1 $B 2>/dev/null | ex "pingpong"
done
SFA_BEV_BAND_PP=1 $B --config density1r 2>/dev/null | ex "pingpong density1r"
SFA_BEV_BAND_PP=1 $B --config argoverse 2>/dev/null | ex "pingpong argoverse"
for pp in 0 1; do for ring in 8 32; do
echo -n "single stream pp=$pp ring=$ring lanes1: "; SFA_BEV_BAND_PP=$pp SFA_BEV_TILED_RING=$ring SFA_BEV_INTERNAL_LANES=1 python tools/bev_run.py 200 3
done; done
