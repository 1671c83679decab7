#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
python lidar*/build.py > /dev/null || exit 1
timeout 900 python -m pytest tests/test_bev_gpu.py tests/test_abi_and_host.py tests/test_bvfeature_gpu.py tests/test_inference_loop_gpu.py tests/test_decode_gpu.py -x -q -m gpu > gpurun_out/r2x_pytest.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/r2x_pytest.log)"
for L in 1 2 3; do echo "single stream, internal lanes $L: $(SFA_BEV_INTERNAL_LANES=$L timeout 120 python tools/bev_run.py 40 3 2>&1 | tail -1)"; done
run() { name=$1; shift
  timeout 600 env "$@" > gpurun_out/r2x_$name.json 2> gpurun_out/r2x_$name.err
  python -c "
import json
try:
    d=json.loads(open('gpurun_out/r2x_$name.json').read().strip().splitlines()[-1])
    print('$name', d['value'], d['ms_per_step'])
except Exception as e: print('$name FAILED', e)"
}
B="python bench.py --steps 1000 --no-e2e --no-cpu-baseline"
for L in 1 2 3; do for p in 1 2 3; do
  run lanes${L}_p$p SFA_BEV_INTERNAL_LANES=$L $B --pipelines $p
done; done
run lanes2_p1_density SFA_BEV_INTERNAL_LANES=2 $B --pipelines 1 --config density1r
run lanes3_p1_density SFA_BEV_INTERNAL_LANES=3 $B --pipelines 1 --config density1r
run lanes2_p2_density SFA_BEV_INTERNAL_LANES=2 $B --pipelines 2 --config density1r
run lanes1_p3_density SFA_BEV_INTERNAL_LANES=1 $B --pipelines 3 --config density1r
