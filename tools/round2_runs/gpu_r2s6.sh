#!/bin/bash
# round 2: (a) fewest band CTAs with the same critical path (512 x 2 items) vs one CTA per slot (592); (b) programmatic dependent
# launch of bev_band behind bev_bin
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_bev_gpu.py -m gpu -x -q > gpurun_out/r2s6_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2s6_pytest.log
SFA_BEV_PDL=1 timeout 600 python -m pytest tests/test_bev_gpu.py tests/test_inference_loop_gpu.py -m gpu -x -q > gpurun_out/r2s6_pytest_pdl.log 2>&1; echo "pytest(pdl) rc=$?"; tail -2 gpurun_out/r2s6_pytest_pdl.log
B="timeout 300 python bench.py --no-e2e --no-cpu-baseline --steps 1500"
ex() { python -c "import json,sys; l=json.loads(sys.stdin.read().strip().splitlines()[-1]); s=l.get('stage_ablation') or {}; k=l['kernels_serialised']; print('$1', round(l['value']), 'single', round((l.get('single_call') or {}).get('value',0)), 'bev_only_us', s.get('bev_only_ms_per_step'), 'band_ms', k['bev_band']['ms_per_step'])"; }
for rep in 1 2; do
$B 2>/dev/null | ex "balanced"
SFA_BAND_FILL_SLOTS=1 $B 2>/dev/null | ex "fill"
SFA_BEV_PDL=1 $B 2>/dev/null | ex "balanced+pdl"
SFA_BEV_PDL=1 SFA_BAND_FILL_SLOTS=1 $B 2>/dev/null | ex "fill+pdl"
done
$B --config density1r 2>/dev/null | ex "balanced density1r"
SFA_BAND_FILL_SLOTS=1 $B --config density1r 2>/dev/null | ex "fill density1r"
SFA_BEV_PDL=1 $B --config density1r 2>/dev/null | ex "balanced+pdl density1r"
$B --config argoverse 2>/dev/null | ex "balanced argoverse"
SFA_BAND_FILL_SLOTS=1 $B --config argoverse 2>/dev/null | ex "fill argoverse"
SFA_BEV_PDL=1 $B --config argoverse 2>/dev/null | ex "balanced+pdl argoverse"
for pdl in 0 1; do echo -n "single stream pdl=$pdl ring=8 lanes1: "; SFA_BEV_PDL=$pdl SFA_BEV_INTERNAL_LANES=1 timeout 120 python tools/bev_run.py 200 3; done
