#!/bin/bash
# round 2: bev_band per-CTA prologue without divisions, no re-zeroing after a CTA's last item
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_bev_gpu.py tests/test_augment_gpu.py tests/test_inference_loop_gpu.py -m gpu -x -q > gpurun_out/r2s5_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2s5_pytest.log
B="python bench.py --no-e2e --no-cpu-baseline --steps 1500"
ex() { python -c "import json,sys; l=json.loads(sys.stdin.read().strip().splitlines()[-1]); s=l.get('stage_ablation') or {}; k=l['kernels_serialised']; print('$1', round(l['value']), 'single', round((l.get('single_call') or {}).get('value',0)), 'bev_only_us', s.get('bev_only_ms_per_step'), 'band_ms', k['bev_band']['ms_per_step'])"; }
for rep in 1 2 3; do $B 2>/dev/null | ex "band-prologue"; done
$B --config density1r 2>/dev/null | ex "density1r"
$B --config argoverse 2>/dev/null | ex "argoverse"
