#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
unset SFA_NVCC_DEFS
python lidar*/build.py > /dev/null || exit 1
timeout 300 python -m pytest tests/test_decode_gpu.py tests/test_inference_loop_gpu.py -x -q -m gpu > gpurun_out/r2m_pytest.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/r2m_pytest.log)"
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum
for ring in 8 16 32; do for eng in 1 3; do
  SFA_BEV_TILED_RING=$ring ncu --replay-mode range --cache-control none --clock-control none --metrics $M --csv --log-file gpurun_out/r2m_range_ring${ring}_e$eng.csv \
     python tools/range_traffic.py 3 6 $eng > gpurun_out/r2m_range_ring${ring}_e$eng.log 2>&1
  echo "range ring$ring engines$eng rc=$? $(grep -v '^==' gpurun_out/r2m_range_ring${ring}_e$eng.csv | tail -3 | cut -d, -f 15 | tr '\n' ' ')"
done; done
run() { name=$1; shift
  timeout 600 env "$@" > gpurun_out/r2m_$name.json 2> gpurun_out/r2m_$name.err
  python -c "
import json
try:
    d=json.loads(open('gpurun_out/r2m_$name.json').read().strip().splitlines()[-1])
    print('$name', d['value'], d['ms_per_step'], {k:round(v['ms_per_step']*1e3,1) for k,v in d['kernels_serialised'].items()})
except Exception as e: print('$name FAILED', e)"
}
B="python bench.py --steps 400 --no-e2e --no-cpu-baseline --lanes 1 --decode-stream"
for ring in 8 15 16 30 32; do for p in 2 3 4; do
  run ring${ring}_p$p SFA_BEV_TILED_RING=$ring $B --pipelines $p
done; done
run ring32_p3_seppost SFA_BEV_TILED_RING=32 $B --pipelines 3 --separate-post
run ring32_p3_nodecstream SFA_BEV_TILED_RING=32 python bench.py --steps 400 --no-e2e --no-cpu-baseline --lanes 1 --pipelines 3
