#!/bin/bash
# round 2: 2-GPU sanity pass of the final tree (bench under torchrun as the driver launches it, the sharded stream)
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 $TR --nproc-per-node 2 --master-port 29612 bench.py --gpus 2 --steps 2000 --warmup 5 > gpurun_out/r2n2_bench_n2.json 2> gpurun_out/r2n2_bench_n2.err; echo "bench n2 rc=$?"
timeout 600 $TR --nproc-per-node 2 --master-port 29613 bench.py --gpus 2 --config stream8192 --no-cpu-baseline > gpurun_out/r2n2_bench_stream_n2.json 2> gpurun_out/r2n2_bench_stream_n2.err; echo "stream n2 rc=$?"
for f in gpurun_out/r2n2_bench_*.json; do echo $f; python -c "
import json,sys
d=json.loads(open('$f').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('value','n_gpus','ms_per_step','steps','warmup')}, 'e2e', (d.get('e2e') or {}).get('value'), (d.get('e2e') or {}).get('frac_of_memcpy_ceiling'))
"; done
