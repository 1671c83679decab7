#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
python lidar*/build.py > /dev/null || exit 1
timeout 900 python -m pytest tests/test_bev_gpu.py tests/test_bvfeature_gpu.py -x -q -m gpu > gpurun_out/r2g_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2g_pytest.log
echo "two-kernel single stream: $(timeout 120 python tools/bev_run.py 30 3 2>&1 | tail -1)"
timeout 300 python tools/bev_distributions.py > gpurun_out/r2g_dist.log 2>&1; grep "us per 64 frames" gpurun_out/r2g_dist.log
timeout 600 python bench.py --steps 400 --no-e2e --no-cpu-baseline --lanes 4 --pipelines 2 > gpurun_out/r2g_bench_l4p2.json 2> gpurun_out/r2g_bench_l4p2.err
timeout 600 python bench.py --steps 400 --no-e2e --no-cpu-baseline > gpurun_out/r2g_bench_l1p1.json 2> gpurun_out/r2g_bench_l1p1.err
timeout 600 python bench.py --steps 400 --no-e2e --no-cpu-baseline --lanes 2 --pipelines 2 > gpurun_out/r2g_bench_l2p2.json 2> gpurun_out/r2g_bench_l2p2.err
for f in gpurun_out/r2g_bench_*.json; do echo $f; python -c "
import json,sys
d=json.loads(open('$f').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], {k:v['ms_per_step'] for k,v in d['kernels_serialised'].items()})
"; done
tail -3 gpurun_out/r2g_bench_l4p2.err
