#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
python __graft_entry__.py build > gpurun_out/r2z_build.log 2>&1 || exit 1
timeout 600 python -m pytest tests/test_bev_gpu.py -x -q -m gpu -k "internal_lanes or concurrent or second_call" > gpurun_out/r2z_pytest.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/r2z_pytest.log)"
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 $TR --nproc-per-node 2 --master-port 29712 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2z_bench_n2.json 2> gpurun_out/r2z_bench_n2.err; echo "bench n2 rc=$?"
timeout 600 $TR --nproc-per-node 2 --master-port 29713 bench.py --gpus 2 --impl reference --steps 2 --warmup 1 > gpurun_out/r2z_ref_n2.json 2> gpurun_out/r2z_ref_n2.err; echo "ref n2 rc=$?"
timeout 600 $TR --nproc-per-node 2 --master-port 29714 bench.py --gpus 2 --config stream8192 --no-cpu-baseline > gpurun_out/r2z_stream_n2.json 2> gpurun_out/r2z_stream_n2.err; echo "stream n2 rc=$?"
for f in gpurun_out/r2z_bench_n2.json gpurun_out/r2z_ref_n2.json gpurun_out/r2z_stream_n2.json; do python -c "
import json
d=json.loads(open('$f').read().strip().splitlines()[-1])
print('$f', {k:d.get(k) for k in ('value','n_gpus','ms_per_step','steps','warmup')}, (d.get('single_call') or {}).get('value'), (d.get('e2e') or {}).get('value'))
"; done
