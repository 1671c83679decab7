#!/bin/bash
# round 2: bev_bin of chunk k+1 launched programmatically behind bev_band of chunk k as well (SFA_BEV_PDL=1; 2 = band only, 0 = off)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2s8_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2s8_pytest.log
B="timeout 300 python bench.py --no-e2e --no-cpu-baseline --steps 1500"
ex() { python -c "import json,sys; l=json.loads(sys.stdin.read().strip().splitlines()[-1]); s=l.get('stage_ablation') or {}; k=l['kernels_serialised']; print('$1', round(l['value']), 'single', round((l.get('single_call') or {}).get('value',0)), 'bev_only_us', s.get('bev_only_ms_per_step'))"; }
for rep in 1 2 3; do
$B 2>/dev/null | ex "pdl both"
SFA_BEV_PDL=2 $B 2>/dev/null | ex "pdl band only"
done
SFA_BEV_PDL=0 $B 2>/dev/null | ex "pdl off"
$B --config density1r 2>/dev/null | ex "pdl both density1r"
$B --config argoverse 2>/dev/null | ex "pdl both argoverse"
for pdl in 0 2 1; do echo -n "single stream pdl=$pdl ring=8 lanes1: "; SFA_BEV_PDL=$pdl SFA_BEV_INTERNAL_LANES=1 timeout 120 python tools/bev_run.py 200 3; done
