#!/bin/bash
# round 2, first GPU pass: parity of the fused BEV kernel + timing against the two-kernel schedule
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_bev_gpu.py -x -q -m gpu > gpurun_out/r2a_pytest_bev.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest_bev.log
tail -5 gpurun_out/r2a_pytest_bev.log
for v in "fused SFA_BEV_FUSED=1" "two_kernel SFA_BEV_FUSED=0" "fused_lag2 SFA_BEV_FUSED_LAG=2" "fused_lag6_ring12 SFA_BEV_FUSED_LAG=6 SFA_BEV_FUSED_RING=12" "fused_lag3_ring6 SFA_BEV_FUSED_LAG=3 SFA_BEV_FUSED_RING=6"; do
  set -- $v; name=$1; shift
  env "$@" timeout 300 python tools/bev_distributions.py > gpurun_out/r2a_dist_$name.log 2>&1; echo "rc=$?" >> gpurun_out/r2a_dist_$name.log
  echo "== $name"; grep "us per 64 frames  (" gpurun_out/r2a_dist_$name.log
done
timeout 600 python bench.py --steps 400 --no-e2e --no-cpu-baseline --lanes 1 --pipelines 1 > gpurun_out/r2a_bench_l1p1.json 2> gpurun_out/r2a_bench_l1p1.err
timeout 600 python bench.py --steps 400 --no-e2e --no-cpu-baseline > gpurun_out/r2a_bench_l4p2.json 2> gpurun_out/r2a_bench_l4p2.err
SFA_BEV_FUSED=0 timeout 600 python bench.py --steps 400 --no-e2e --no-cpu-baseline > gpurun_out/r2a_bench_twokernel.json 2> gpurun_out/r2a_bench_twokernel.err
for f in gpurun_out/r2a_bench_*.json; do echo $f; python -c "
import json,sys
d=json.loads(open('$f').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], {k:v['ms_per_step'] for k,v in d['kernels'].items()})
"; done
