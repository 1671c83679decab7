#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
run() { # name, env/args...
  name=$1; shift
  timeout 600 env "$@" > gpurun_out/r2l_$name.json 2> gpurun_out/r2l_$name.err
  python -c "
import json
try:
    d=json.loads(open('gpurun_out/r2l_$name.json').read().strip().splitlines()[-1])
    print('$name', d['value'], d['ms_per_step'], {k:round(v['ms_per_step']*1e3,1) for k,v in d['kernels_serialised'].items()})
except Exception as e: print('$name FAILED', e)"
}
B="python bench.py --steps 400 --no-e2e --no-cpu-baseline"
for v in "sel256l2 -DSFA_SEL_THREADS=256 -DSFA_SEL_SMEM_ITEMS=1024" "sel128l2 -DSFA_SEL_THREADS=128 -DSFA_SEL_SMEM_ITEMS=1024" "sel512l2 -DSFA_SEL_THREADS=512 -DSFA_SEL_SMEM_ITEMS=1024" "sel1024 -DSFA_SEL_THREADS=1024"; do
  set -- $v; name=$1; shift
  export SFA_NVCC_DEFS="$*"
  python lidar*/build.py > /dev/null || { echo "build failed $name"; continue; }
  timeout 300 python -m pytest tests/test_decode_gpu.py -x -q -m gpu > gpurun_out/r2l_pytest_$name.log 2>&1; echo "$name pytest rc=$? $(tail -1 gpurun_out/r2l_pytest_$name.log)"
  run ${name}_l2p2 $B
  run ${name}_l2p2_lowprio $B --decode-low-priority 1
done
# ring size x schedule with the small select
export SFA_NVCC_DEFS="-DSFA_SEL_THREADS=256 -DSFA_SEL_SMEM_ITEMS=1024"
python lidar*/build.py > /dev/null
for ring in 8 16 32; do
  run ring${ring}_l2p2 SFA_BEV_TILED_RING=$ring $B
  run ring${ring}_l4p2 SFA_BEV_TILED_RING=$ring $B --lanes 4
  run ring${ring}_l1p2 SFA_BEV_TILED_RING=$ring $B --lanes 1 --decode-stream
  run ring${ring}_l1p3 SFA_BEV_TILED_RING=$ring $B --lanes 1 --decode-stream --pipelines 3
done
