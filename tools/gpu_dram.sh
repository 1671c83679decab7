#!/bin/bash
# Single-pass ncu collection (no kernel replay, caches untouched) of DRAM bytes per launch for an eager step.
set -u
CMD="python bench.py --steps 2 --warmup 3 --eager --lanes 1 --no-e2e --no-cpu-baseline"
mkdir -p gpurun_out
for RING in ${RINGS:-32 16 8}; do
  SFA_BEV_TILED_RING=$RING ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct \
      --clock-control none --cache-control none -k regex:"bev_|decode" -s 40 -c 24 --csv --log-file gpurun_out/dram_ring$RING.csv $CMD > gpurun_out/dram_ring$RING.log 2>&1
  echo "ring $RING rc=$?"
done
