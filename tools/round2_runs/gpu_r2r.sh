#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
python lidar*/build.py > /dev/null || exit 1
run() { name=$1; shift
  timeout 600 env "$@" > gpurun_out/r2r_$name.json 2> gpurun_out/r2r_$name.err
  python -c "
import json
try:
    d=json.loads(open('gpurun_out/r2r_$name.json').read().strip().splitlines()[-1])
    print('$name', d['value'], d['ms_per_step'])
except Exception as e: print('$name FAILED', e)"
}
B="python bench.py --steps 1500 --no-e2e --no-cpu-baseline"
for rep in 1 2 3; do
  run p2_$rep $B --pipelines 2
  run p3_$rep $B --pipelines 3
  run r16p2_$rep SFA_BEV_TILED_RING=16 $B --pipelines 2
  run r12p2_$rep SFA_BEV_TILED_RING=12 $B --pipelines 2
done
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum
for v in "8 2" "8 3" "12 2"; do set -- $v
  SFA_BEV_TILED_RING=$1 ncu --replay-mode range --cache-control none --clock-control none --metrics $M --csv --log-file gpurun_out/r2r_range_ring$1_e$2.csv \
     python tools/range_traffic.py 3 6 $2 > gpurun_out/r2r_range_ring$1_e$2.log 2>&1
  echo "range ring$1 engines$2: $(grep -v '^==' gpurun_out/r2r_range_ring$1_e$2.csv | tail -3 | awk -F'\",\"' '{print $(NF-2), $NF}' | tr '\n' ' ')"
done
