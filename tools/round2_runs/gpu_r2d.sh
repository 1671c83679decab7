#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
export SFA_NVCC_DEFS="-DSFA_FUSED_WORKERS=224"
python lidar*/build.py > /dev/null || exit 1
for v in "lag6r12 SFA_BEV_FUSED_LAG=6 SFA_BEV_FUSED_RING=12" "lag8r12 SFA_BEV_FUSED_LAG=8 SFA_BEV_FUSED_RING=12" "lag10r16 SFA_BEV_FUSED_LAG=10 SFA_BEV_FUSED_RING=16" "lag14r20 SFA_BEV_FUSED_LAG=14 SFA_BEV_FUSED_RING=20" "lag20r28 SFA_BEV_FUSED_LAG=20 SFA_BEV_FUSED_RING=28"; do
  set -- $v; name=$1; shift
  echo "$name: $(env "$@" timeout 120 python tools/bev_run.py 30 1 2>&1 | tail -1)"
done
export SFA_BEV_FUSED_LAG=8 SFA_BEV_FUSED_RING=12
ncu --set full --clock-control none --cache-control none --import-source on -k regex:bev_fused -s 4 -c 1 -f -o gpurun_out/r2d_fused python tools/bev_run.py 4 1 > gpurun_out/r2d_ncu.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/r2d_ncu.log
