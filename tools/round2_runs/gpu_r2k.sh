#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
# ---- fused kernel shapes ----
for v in "w224c4 -DSFA_FUSED_WORKERS=224 -DSFA_FUSED_CTAS=4 -DSFA_FUSED_NB=128" "w448c2 -DSFA_FUSED_WORKERS=448 -DSFA_FUSED_CTAS=2 -DSFA_FUSED_NB=64" "w320c3 -DSFA_FUSED_WORKERS=320 -DSFA_FUSED_CTAS=3 -DSFA_FUSED_NB=96"; do
  set -- $v; name=$1; shift
  export SFA_NVCC_DEFS="$*"
  python lidar*/build.py > /dev/null || { echo "build failed $name"; continue; }
  timeout 600 python -m pytest tests/test_bev_gpu.py -x -q -m gpu -k "tiled" > gpurun_out/r2k_pytest_$name.log 2>&1; echo "$name pytest rc=$? $(tail -1 gpurun_out/r2k_pytest_$name.log)"
  for l in "6 10" "8 12" "10 16" "14 20"; do
    set -- $l
    echo "$name lag$1 ring$2: $(SFA_BEV_FUSED_LAG=$1 SFA_BEV_FUSED_RING=$2 timeout 120 python tools/bev_run.py 30 1 2>&1 | tail -1)"
  done
done
unset SFA_NVCC_DEFS
python lidar*/build.py > /dev/null
# ---- whole-step DRAM traffic, range replay ----
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum
for v in "twokernel_ring32 3 SFA_BEV_TILED_RING=32" "twokernel_ring8 3 SFA_BEV_TILED_RING=8" "fused 1 SFA_X=1"; do
  set -- $v; name=$1; algo=$2; shift 2
  env "$@" ncu --replay-mode range --cache-control none --clock-control none --metrics $M --csv --log-file gpurun_out/r2k_range_$name.csv \
     python tools/range_traffic.py $algo 2 > gpurun_out/r2k_range_$name.log 2>&1
  echo "range $name rc=$?"; grep -v "^==" gpurun_out/r2k_range_$name.csv | tail -4 | cut -d, -f 5,12-15; tail -2 gpurun_out/r2k_range_$name.log
done
# ---- decode select shapes ----
for v in "sel1024 -DSFA_SEL_THREADS=1024" "sel512 -DSFA_SEL_THREADS=512" "sel256 -DSFA_SEL_THREADS=256" "sel256l2 -DSFA_SEL_THREADS=256 -DSFA_SEL_SMEM_ITEMS=1024"; do
  set -- $v; name=$1; shift
  export SFA_NVCC_DEFS="$*"
  python lidar*/build.py > /dev/null || { echo "build failed $name"; continue; }
  timeout 300 python -m pytest tests/test_decode_gpu.py -x -q -m gpu > gpurun_out/r2k_pytest_$name.log 2>&1; echo "$name pytest rc=$? $(tail -1 gpurun_out/r2k_pytest_$name.log)"
  timeout 600 python bench.py --steps 400 --no-e2e --no-cpu-baseline > gpurun_out/r2k_bench_$name.json 2> gpurun_out/r2k_bench_$name.err
  python -c "
import json
d=json.loads(open('gpurun_out/r2k_bench_$name.json').read().strip().splitlines()[-1])
print('$name', d['value'], d['ms_per_step'], {k:v['ms_per_step'] for k,v in d['kernels_serialised'].items()})"
done
