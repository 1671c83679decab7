#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
python lidar*/build.py > /dev/null || exit 1
timeout 900 python -m pytest tests/test_bev_gpu.py tests/test_bvfeature_gpu.py -x -q -m gpu > gpurun_out/r2o_pytest.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/r2o_pytest.log)"
timeout 900 python -m pytest tests/test_bev_gpu.py -x -q -m gpu -k "two_kernel or overflow or crowded or front_back or extras" > gpurun_out/r2o_pytest2.log 2>&1; echo "pytest(2nd pass) rc=$? $(tail -1 gpurun_out/r2o_pytest2.log)"
echo "two-kernel single stream ring8: $(timeout 120 python tools/bev_run.py 40 3 2>&1 | tail -1)"
echo "two-kernel single stream ring32: $(SFA_BEV_TILED_RING=32 timeout 120 python tools/bev_run.py 40 3 2>&1 | tail -1)"
run() { name=$1; shift
  timeout 600 env "$@" > gpurun_out/r2o_$name.json 2> gpurun_out/r2o_$name.err
  python -c "
import json
try:
    d=json.loads(open('gpurun_out/r2o_$name.json').read().strip().splitlines()[-1])
    print('$name', d['value'], d['ms_per_step'], {k:round(v['ms_per_step']*1e3,1) for k,v in d['kernels_serialised'].items()})
except Exception as e: print('$name FAILED', e)"
}
B="python bench.py --steps 600 --no-e2e --no-cpu-baseline"
run ring8_p3 $B
run ring8_p4 $B --pipelines 4
run ring16_p2 SFA_BEV_TILED_RING=16 $B --pipelines 2
run ring16_p3 SFA_BEV_TILED_RING=16 $B --pipelines 3
run ring32_p3 SFA_BEV_TILED_RING=32 $B --pipelines 3
