#!/bin/bash
# DRAM bytes of one whole step (one CUDA-graph launch = 64 frames of BEV + decode) in the benchmarked schedule
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
unset SFA_NVCC_DEFS
python lidar*/build.py > /dev/null || exit 1
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct
for v in "l2p2_ring32 2 2 SFA_BEV_TILED_RING=32" "l2p2_ring8 2 2 SFA_BEV_TILED_RING=8" "l1p1_ring32 1 1 SFA_BEV_TILED_RING=32" "l1p1_ring8 1 1 SFA_BEV_TILED_RING=8" "l1p1_fused 1 1 SFA_BEV_FUSED=1"; do
  set -- $v; name=$1; lanes=$2; pipes=$3; shift 3
  env "$@" ncu --graph-profiling graph --cache-control none --clock-control none --metrics $M --launch-skip 4 -c 6 --csv --log-file gpurun_out/r2i_graph_$name.csv \
     python bench.py --steps 8 --warmup 3 --settle-s 0 --no-e2e --no-cpu-baseline --lanes $lanes --pipelines $pipes > gpurun_out/r2i_$name.log 2>&1
  echo "$name rc=$?"
  python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2i_graph_$name.csv')) if len(r)>10]
hdr=rows[0]; mi=hdr.index('Metric Name'); vi=hdr.index('Metric Value'); ii=hdr.index('ID')
agg={}
for r in rows[1:]:
    agg.setdefault(r[ii],{})[r[mi]]=float(r[vi].replace(',',''))
for k,v in agg.items():
    print('$name', k, {m:round(x/1e6,1) if 'bytes' in m else round(x,1) for m,x in v.items()})
PY
done
