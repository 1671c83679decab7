#!/bin/bash
# (on the GPU box) ncu with --graph-profiling graph: every CUDA-graph replay of the bench (a whole step: 4 BEV lanes +
# decode running CONCURRENTLY) is one profiled unit -> real DRAM traffic and pipe utilisation of a step.
set -u
CMD="python bench.py --steps 4 --warmup 3 --pipelines 1 --no-e2e --no-cpu-baseline"
mkdir -p gpurun_out
$CMD > gpurun_out/plain_graph.log 2>&1 || { echo "plain run failed"; tail -20 gpurun_out/plain_graph.log; exit 1; }
ncu --graph-profiling graph --clock-control none --cache-control none \
    --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,lts__throughput.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_elapsed,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,l1tex__data_pipe_lsu_wavefronts_mem_shared.avg.pct_of_peak_sustained_elapsed,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active \
    -s 10 -c 6 --csv --log-file gpurun_out/graph_profile.csv $CMD > gpurun_out/ncu_graph.log 2>&1
echo "graph profile rc=$?"; tail -3 gpurun_out/ncu_graph.log
