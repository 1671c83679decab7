#!/bin/bash
# Runs on the GPU box (under gpurun): plain run, then (optionally) the ncu launch list and one full
# capture of the kernels matching $1 (regex) of the same command.  Outputs land in gpurun_out/.
#   tools/gpu_profile.sh 'decode_kernel|bev_band_kernel|bev_bin_kernel' 9 [list]
set -u
PAT="${1:-decode_kernel|bev_band_kernel|bev_bin_kernel}"
COUNT="${2:-9}"
CMD="python bench.py --steps 2 --warmup 3 --eager --lanes 1 --no-e2e --no-cpu-baseline"
mkdir -p gpurun_out
$CMD > gpurun_out/plain.log 2>&1 || { echo "plain run failed"; tail -20 gpurun_out/plain.log; exit 1; }
if [ "${3:-}" = "list" ]; then
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
  echo "launch list rc=$?"
fi
ncu --set full --clock-control none --cache-control ${CACHE:-all} --import-source on -k regex:"$PAT" -s "${SKIP:-24}" -c "$COUNT" -f -o gpurun_out/prof $CMD > gpurun_out/ncu_full.log 2>&1
echo "full capture rc=$?"
tail -2 gpurun_out/ncu_full.log
