#!/bin/bash
# 8-GPU box: memcpy-only ceiling of the e2e leg at 1/2/4/8 GPUs, the bench at 8 and 2 GPUs, the sharded 8192-frame stream
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
python __graft_entry__.py build > gpurun_out/r2s_build.log 2>&1 || exit 1
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 200 python tools/pcie_ceiling.py > gpurun_out/r2s_pcie_n1.json 2> gpurun_out/r2s_pcie_n1.err; tail -1 gpurun_out/r2s_pcie_n1.json
for n in 2 4 8; do
  timeout 300 $TR --nproc-per-node $n --master-port $((29500+n)) tools/pcie_ceiling.py > gpurun_out/r2s_pcie_n$n.json 2> gpurun_out/r2s_pcie_n$n.err; tail -1 gpurun_out/r2s_pcie_n$n.json
done
timeout 600 $TR --nproc-per-node 8 --master-port 29611 bench.py --gpus 8 --steps 2000 --warmup 5 > gpurun_out/r2s_bench_n8.json 2> gpurun_out/r2s_bench_n8.err; echo "bench n8 rc=$?"
timeout 600 $TR --nproc-per-node 2 --master-port 29612 bench.py --gpus 2 --steps 2000 --warmup 5 > gpurun_out/r2s_bench_n2.json 2> gpurun_out/r2s_bench_n2.err; echo "bench n2 rc=$?"
timeout 600 $TR --nproc-per-node 8 --master-port 29613 bench.py --gpus 8 --config stream8192 --no-cpu-baseline > gpurun_out/r2s_bench_stream_n8.json 2> gpurun_out/r2s_bench_stream_n8.err; echo "stream n8 rc=$?"
timeout 600 $TR --nproc-per-node 8 --master-port 29614 bench.py --gpus 8 --impl reference --steps 2 --warmup 1 > gpurun_out/r2s_bench_ref_n8.json 2> gpurun_out/r2s_bench_ref_n8.err; echo "ref n8 rc=$?"
for f in gpurun_out/r2s_bench_*.json; do echo $f; python -c "
import json,sys
d=json.loads(open('$f').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('value','n_gpus','ms_per_step','steps','warmup')}, 'e2e', (d.get('e2e') or {}), (d.get('config') or {}).get('stream'))
"; done
