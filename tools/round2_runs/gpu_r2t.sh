#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
python lidar*/build.py > /dev/null || exit 1
timeout 600 python -m pytest tests/test_bev_gpu.py -x -q -m gpu -k "two_kernel or batch64 or uniform or full_size" > gpurun_out/r2t_pytest.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/r2t_pytest.log)"
run() { name=$1; shift
  timeout 600 env "$@" > gpurun_out/r2t_$name.json 2> gpurun_out/r2t_$name.err
  python -c "
import json
try:
    d=json.loads(open('gpurun_out/r2t_$name.json').read().strip().splitlines()[-1])
    print('$name', d['value'], d['ms_per_step'])
except Exception as e: print('$name FAILED', e)"
}
B="python bench.py --steps 1500 --no-e2e --no-cpu-baseline"
for rep in 1 2 3; do
  run pf1_$rep SFA_BEV_PREFETCH=1 $B
  run pf0_$rep SFA_BEV_PREFETCH=0 $B
done
run pf1_r16 SFA_BEV_PREFETCH=1 SFA_BEV_TILED_RING=16 $B
run pf0_r16 SFA_BEV_PREFETCH=0 SFA_BEV_TILED_RING=16 $B
echo "single stream pf1: $(SFA_BEV_PREFETCH=1 timeout 120 python tools/bev_run.py 40 3 2>&1 | tail -1)"
echo "single stream pf0: $(SFA_BEV_PREFETCH=0 timeout 120 python tools/bev_run.py 40 3 2>&1 | tail -1)"
