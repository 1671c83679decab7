#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
python lidar*/build.py > /dev/null || exit 1
timeout 900 python -m pytest tests/test_bev_gpu.py tests/test_bvfeature_gpu.py tests/test_abi_and_host.py -x -q -m gpu > gpurun_out/r2q_pytest.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/r2q_pytest.log)"
run() { name=$1; shift
  timeout 600 env "$@" > gpurun_out/r2q_$name.json 2> gpurun_out/r2q_$name.err
  python -c "
import json
try:
    d=json.loads(open('gpurun_out/r2q_$name.json').read().strip().splitlines()[-1])
    print('$name', d['value'], d['ms_per_step'], d['gpu_launches'], {k:round(v['ms_per_step']*1e3,1) for k,v in d['kernels_serialised'].items()})
except Exception as e: print('$name FAILED', e)"
}
B="python bench.py --steps 800 --no-e2e --no-cpu-baseline"
run ring8_p3_a $B
run ring8_p3_b $B
run ring8_p2 $B --pipelines 2
run ring8_p4 $B --pipelines 4
echo "single stream ring8: $(timeout 120 python tools/bev_run.py 40 3 2>&1 | tail -1)"
