#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
run() { name=$1; shift
  timeout 600 env "$@" > gpurun_out/r2w_$name.json 2> gpurun_out/r2w_$name.err
  python -c "
import json
try:
    d=json.loads(open('gpurun_out/r2w_$name.json').read().strip().splitlines()[-1])
    print('$name', d['value'], d['ms_per_step'], {k:round(v['ms_per_step']*1e3,1) for k,v in d['kernels_serialised'].items()})
except Exception as e: print('$name FAILED', e)"
}
for rows in 19 16 17 20; do
  export SFA_NVCC_DEFS="-DSFA_CAND_ROWS=$rows"
  python lidar*/build.py > /dev/null || { echo "build failed"; continue; }
  timeout 600 python -m pytest tests/test_decode_gpu.py tests/test_inference_loop_gpu.py -x -q -m gpu > gpurun_out/r2w_pytest_$rows.log 2>&1; echo "rows=$rows pytest rc=$? $(tail -1 gpurun_out/r2w_pytest_$rows.log)"
  run rows${rows}_a python bench.py --steps 1200 --no-e2e --no-cpu-baseline
  run rows${rows}_b python bench.py --steps 1200 --no-e2e --no-cpu-baseline
  run rows${rows}_deconly python bench.py --steps 1200 --no-e2e --no-cpu-baseline --only decode
done
