#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
python lidar*/build.py > /dev/null || exit 1
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum
for v in "16 2" "12 2" "12 3" "10 3" "8 3"; do set -- $v
  SFA_BEV_TILED_RING=$1 ncu --replay-mode range --cache-control none --clock-control none --metrics $M --csv --log-file gpurun_out/r2p_range_ring$1_e$2.csv \
     python tools/range_traffic.py 3 6 $2 > gpurun_out/r2p_range_ring$1_e$2.log 2>&1
  echo "range ring$1 engines$2: $(grep -v '^==' gpurun_out/r2p_range_ring$1_e$2.csv | tail -3 | awk -F'\",\"' '{print $(NF-2), $NF}' | tr '\n' ' ')"
done
run() { name=$1; shift
  timeout 600 env "$@" > gpurun_out/r2p_$name.json 2> gpurun_out/r2p_$name.err
  python -c "
import json
try:
    d=json.loads(open('gpurun_out/r2p_$name.json').read().strip().splitlines()[-1])
    print('$name', d['value'], d['ms_per_step'], {k:round(v['ms_per_step']*1e3,1) for k,v in d['kernels_serialised'].items()})
except Exception as e: print('$name FAILED', e)"
}
B="python bench.py --steps 800 --no-e2e --no-cpu-baseline"
run ring16_p2_a SFA_BEV_TILED_RING=16 $B --pipelines 2
run ring16_p2_b SFA_BEV_TILED_RING=16 $B --pipelines 2
run ring12_p2 SFA_BEV_TILED_RING=12 $B --pipelines 2
run ring12_p3 SFA_BEV_TILED_RING=12 $B --pipelines 3
run ring8_p3_a $B
run ring8_p3_b $B
run ring8_p3_bevonly $B --only bev
run ring8_p3_deconly $B --only decode
run ring16_p2_bevonly SFA_BEV_TILED_RING=16 $B --pipelines 2 --only bev
