#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
for v in "base" "nozero -DSFA_BAND_TMA_ZERO=0" "ldgnc -DSFA_RECORD_LDG_NC" "both -DSFA_BAND_TMA_ZERO=0 -DSFA_RECORD_LDG_NC"; do
  set -- $v; name=$1; shift
  export SFA_NVCC_DEFS="$*"
  python lidar*/build.py > /dev/null || exit 1
  echo "$name single stream: $(timeout 120 python tools/bev_run.py 40 3 2>&1 | tail -1)"
  timeout 600 python bench.py --steps 400 --no-e2e --no-cpu-baseline --lanes 2 --pipelines 2 > gpurun_out/r2h_bench_$name.json 2> gpurun_out/r2h_bench_$name.err
  python -c "
import json
d=json.loads(open('gpurun_out/r2h_bench_$name.json').read().strip().splitlines()[-1])
print('$name l2p2', d['value'], d['ms_per_step'], {k:v['ms_per_step'] for k,v in d['kernels_serialised'].items()})"
done
