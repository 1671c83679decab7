#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
python tools/bev_run.py 20 > gpurun_out/r2b_plain.log 2>&1 || { tail -20 gpurun_out/r2b_plain.log; exit 1; }
cat gpurun_out/r2b_plain.log
ncu --set full --clock-control none --cache-control none --import-source on -k regex:bev_fused -s 4 -c 1 -f -o gpurun_out/r2b_fused python tools/bev_run.py 4 > gpurun_out/r2b_ncu.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/r2b_ncu.log
