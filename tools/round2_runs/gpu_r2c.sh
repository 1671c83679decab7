#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
for W in 256 224; do
  export SFA_NVCC_DEFS="-DSFA_FUSED_WORKERS=$W"
  python lidar*/build.py > /dev/null || exit 1
  timeout 600 python -m pytest tests/test_bev_gpu.py -x -q -m gpu -k "tiled" > gpurun_out/r2c_pytest_$W.log 2>&1; echo "W=$W pytest rc=$?"; tail -2 gpurun_out/r2c_pytest_$W.log
  for v in "d SFA_X=0" "lag6 SFA_BEV_FUSED_LAG=6 SFA_BEV_FUSED_RING=12" "lag3 SFA_BEV_FUSED_LAG=3 SFA_BEV_FUSED_RING=6" "lag2 SFA_BEV_FUSED_LAG=2 SFA_BEV_FUSED_RING=4"; do
    set -- $v; name=$1; shift
    echo "W=$W $name: $(env "$@" timeout 120 python tools/bev_run.py 30 1 2>&1 | tail -1)"
  done
  SFA_BEV_FUSED=1 timeout 300 python tools/bev_distributions.py > gpurun_out/r2c_dist_$W.log 2>&1; grep "us per 64 frames  (" gpurun_out/r2c_dist_$W.log
done
echo "two-kernel: $(timeout 120 python tools/bev_run.py 30 3 2>&1 | tail -1)"
