#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
python lidar*/build.py > /dev/null || exit 1
run() { name=$1; shift
  timeout 600 env "$@" > gpurun_out/r2u_$name.json 2> gpurun_out/r2u_$name.err
  python -c "
import json
try:
    d=json.loads(open('gpurun_out/r2u_$name.json').read().strip().splitlines()[-1])
    print('$name', d['value'], d['ms_per_step'])
except Exception as e: print('$name FAILED', e)"
}
for cfg in density1r argoverse headline; do for p in 2 3 4; do for rep in 1 2; do
  run ${cfg}_p${p}_$rep python bench.py --config $cfg --steps 800 --no-e2e --no-cpu-baseline --pipelines $p
done; done; done
