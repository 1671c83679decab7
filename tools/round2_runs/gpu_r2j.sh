#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
unset SFA_NVCC_DEFS
python lidar*/build.py > /dev/null || exit 1
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2j_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2j_pytest.log
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum
for v in "twokernel_ring32 3 SFA_BEV_TILED_RING=32" "twokernel_ring8 3 SFA_BEV_TILED_RING=8" "fused 1 SFA_X=1"; do
  set -- $v; name=$1; algo=$2; shift 2
  env "$@" ncu --replay-mode range --profile-from-start off --cache-control none --clock-control none --metrics $M --csv --log-file gpurun_out/r2j_range_$name.csv \
     python tools/range_traffic.py $algo 2 > gpurun_out/r2j_range_$name.log 2>&1
  echo "range $name rc=$?"; grep -v "^==" gpurun_out/r2j_range_$name.csv | tail -4 | cut -d, -f 5,13-15 
done
for cfg in headline density1r argoverse loop32; do
  timeout 600 python bench.py --config $cfg --steps 300 > gpurun_out/r2j_bench_$cfg.json 2> gpurun_out/r2j_bench_$cfg.err; echo "bench $cfg rc=$?"
done
timeout 900 python bench.py --config stream8192 --no-cpu-baseline > gpurun_out/r2j_bench_stream8192.json 2> gpurun_out/r2j_bench_stream8192.err; echo "bench stream rc=$?"
timeout 300 python tools/pcie_ceiling.py > gpurun_out/r2j_pcie_n1.json 2>&1
for f in gpurun_out/r2j_bench_*.json; do echo $f; python -c "
import json,sys
d=json.loads(open('$f').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('value','ms_per_step','steps','warmup')}, (d.get('e2e') or {}).get('value'), d.get('roofline'), d.get('loop'), (d.get('cpu_baseline') or {}).get('value'))
"; done
cat gpurun_out/r2j_pcie_n1.json | tail -1
