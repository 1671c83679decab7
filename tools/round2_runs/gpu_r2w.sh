#!/bin/bash
# round 2: wave-fitting experiments — bev_bin tile size (one wave per launch) and 148 bands (two band items per CTA)
mkdir -p gpurun_out
python -m pytest tests/test_bev_gpu.py -m gpu -x -q > gpurun_out/r2w_pytest.log 2>&1; echo "pytest(auto tile,128) rc=$?"; tail -2 gpurun_out/r2w_pytest.log
SFA_BEV_BANDS=148 python -m pytest tests/test_bev_gpu.py tests/test_augment_gpu.py -m gpu -x -q > gpurun_out/r2w_pytest148.log 2>&1; echo "pytest(148) rc=$?"; tail -2 gpurun_out/r2w_pytest148.log
B="python bench.py --no-e2e --no-cpu-baseline --steps 1500"
ex() { python -c "import json,sys; l=json.loads(sys.stdin.read().strip().splitlines()[-1]); s=l.get('stage_ablation') or {}; print('$1', round(l['value']), 'single', round((l.get('single_call') or {}).get('value',0)), 'bev_only_us', s.get('bev_only_ms_per_step'))"; }
for rep in 1 2; do
SFA_BIN_TILE=2048 $B 2>/dev/null | ex "tile2048 bands128"
$B 2>/dev/null | ex "tileauto bands128"
SFA_BIN_TILE=2048 SFA_BEV_BANDS=148 $B 2>/dev/null | ex "tile2048 bands148"
SFA_BEV_BANDS=148 $B 2>/dev/null | ex "tileauto bands148"
done
for t in 2048 0; do for nb in 128 148; do
echo -n "single stream tile=$t bands=$nb lanes1: "; SFA_BEV_INTERNAL_LANES=1 SFA_BIN_TILE=$t SFA_BEV_BANDS=$nb python tools/bev_run.py 200 3
echo -n "single stream tile=$t bands=$nb lanes2: "; SFA_BIN_TILE=$t SFA_BEV_BANDS=$nb python tools/bev_run.py 200 3
done; done
