#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
unset SFA_NVCC_DEFS
python __graft_entry__.py build > gpurun_out/r2f_build.log 2>&1 || exit 1
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2f_pytest.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/r2f_pytest.log)"
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2f_smoke.log 2>&1; echo "smoke rc=$? $(tail -1 gpurun_out/r2f_smoke.log)"
timeout 900 python bench.py > gpurun_out/r2f_bench_headline.json 2> gpurun_out/r2f_bench_headline.err; echo "bench rc=$?"
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2f_bench_headline_driver_args.json 2> gpurun_out/r2f_bench_headline_driver_args.err; echo "bench(driver args) rc=$?"
timeout 600 python bench.py --impl reference --steps 4 --warmup 1 > gpurun_out/r2f_bench_reference.json 2> gpurun_out/r2f_bench_reference.err; echo "reference rc=$?"
for cfg in density1r argoverse loop32; do
  timeout 600 python bench.py --config $cfg --steps 400 > gpurun_out/r2f_bench_$cfg.json 2> gpurun_out/r2f_bench_$cfg.err; echo "bench $cfg rc=$?"
done
timeout 900 python bench.py --config stream8192 --no-cpu-baseline > gpurun_out/r2f_bench_stream8192.json 2> gpurun_out/r2f_bench_stream8192.err; echo "bench stream rc=$?"
for f in gpurun_out/r2f_bench_*.json; do echo $f; python -c "
import json,sys
d=json.loads(open('$f').read().strip().splitlines()[-1])
r=d.get('roofline') or {}
print({k:d.get(k) for k in ('value','ms_per_step','steps','warmup')}, 'e2e', (d.get('e2e') or {}).get('value'), 'roof', r.get('kernel'), r.get('frac'), r.get('traffic'), d.get('loop'), 'cpu', (d.get('cpu_baseline') or {}).get('value'), (d.get('cpu_baseline') or {}).get('kind'), d.get('clocks'))
"; done
# ncu: launch list, then a full capture of the hot kernels, of the same (eager, serialised) command
CMD="python bench.py --eager --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --pipelines 1"
$CMD > gpurun_out/r2f_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/r2f_plain.log; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2f_launches.csv $CMD > gpurun_out/r2f_ncu_list.log 2>&1; echo "launch list rc=$?"
ncu --set full --clock-control none --cache-control none --import-source on -k regex:"bev_bin_staged_kernel|bev_band_kernel|peak_candidates_kernel|peak_select_kernel" -s 64 -c 8 -f -o gpurun_out/r2f_prof $CMD > gpurun_out/r2f_ncu_full.log 2>&1; echo "full capture rc=$?"
