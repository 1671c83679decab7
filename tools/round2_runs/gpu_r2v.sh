#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
python lidar*/build.py > /dev/null || exit 1
timeout 900 python -m pytest tests/test_bev_gpu.py -x -q -m gpu > gpurun_out/r2v_pytest.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/r2v_pytest.log)"
run() { name=$1; shift
  timeout 600 env "$@" > gpurun_out/r2v_$name.json 2> gpurun_out/r2v_$name.err
  python -c "
import json
try:
    d=json.loads(open('gpurun_out/r2v_$name.json').read().strip().splitlines()[-1])
    print('$name', d['value'], d['ms_per_step'], {k:round(v['ms_per_step']*1e3,1) for k,v in d['kernels_serialised'].items()})
except Exception as e: print('$name FAILED', e)"
}
for p in 2 3; do for rep in 1 2; do
  run argoverse_p${p}_$rep python bench.py --config argoverse --steps 600 --no-e2e --no-cpu-baseline --pipelines $p
done; done
run density1r_p3 python bench.py --config density1r --steps 600 --no-e2e --no-cpu-baseline
run headline_p3 python bench.py --steps 600 --no-e2e --no-cpu-baseline
SFA_N=250000 timeout 120 python tools/bev_run.py 20 3 2>&1 | tail -1
