#!/bin/bash
# round 2: peak_candidates staged by TMA bulk loads + in-place key conversion
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2s4_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2s4_pytest.log
B="python bench.py --no-e2e --no-cpu-baseline --steps 1500"
ex() { python -c "import json,sys; l=json.loads(sys.stdin.read().strip().splitlines()[-1]); s=l.get('stage_ablation') or {}; k=l['kernels_serialised']; print('$1', round(l['value']), 'single', round((l.get('single_call') or {}).get('value',0)), 'bev_only_us', s.get('bev_only_ms_per_step'), 'dec_only', s.get('decode_only_ms_per_step'), 'adds', s.get('decode_adds_ms_per_step'), 'cand_ms', k['peak_candidates']['ms_per_step'], 'sel_ms', k['peak_select']['ms_per_step'])"; }
for rep in 1 2 3; do $B 2>/dev/null | ex "cand-tma"; done
python tools/decode_realistic.py 2>&1 | tail -8
