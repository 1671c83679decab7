#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
unset SFA_NVCC_DEFS
python lidar*/build.py > /dev/null || exit 1
run() { name=$1; shift
  timeout 600 env "$@" > gpurun_out/r2n_$name.json 2> gpurun_out/r2n_$name.err
  python -c "
import json
try:
    d=json.loads(open('gpurun_out/r2n_$name.json').read().strip().splitlines()[-1])
    print('$name', d['value'], d['ms_per_step'], {k:round(v['ms_per_step']*1e3,1) for k,v in d['kernels_serialised'].items()})
except Exception as e: print('$name FAILED', e)"
}
B="python bench.py --steps 600 --no-e2e --no-cpu-baseline --lanes 1 --decode-stream"
for ring in 6 7 8 9 10 12; do for p in 3 4 5; do
  run ring${ring}_p$p SFA_BEV_TILED_RING=$ring $B --pipelines $p
done; done
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum
for ring in 7 10 12; do
  SFA_BEV_TILED_RING=$ring ncu --replay-mode range --cache-control none --clock-control none --metrics $M --csv --log-file gpurun_out/r2n_range_ring${ring}_e4.csv \
     python tools/range_traffic.py 3 8 4 > gpurun_out/r2n_range_ring${ring}_e4.log 2>&1
  echo "range ring$ring engines4: $(grep -v '^==' gpurun_out/r2n_range_ring${ring}_e4.csv | tail -3 | awk -F'\",\"' '{print $(NF-2), $NF}' | tr '\n' ' ')"
done
