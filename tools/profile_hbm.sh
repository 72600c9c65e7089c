#!/bin/bash
# GPU-box profiling pass (run under gpurun; small CSV outputs only -- a full-set .ncu-rep of a whole step exceeds
# gpurun's 64 MiB return limit):
#   1. every kernel of ONE eager CelebA step at B=1024 with its time and DRAM bytes  -> <tag>_step_kernels.csv
#   2. tensor-pipe activity of the nine big-layer launches (tools/bench_layers.py)   -> <tag>_tensor_<i>.csv
# usage: tools/profile_hbm.sh <tag>   ; summarise with tools/ncu_hbm_summarize.py <tag>
set -u
tag=${1:-rXX}
mkdir -p gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,launch__grid_size
CMD="python bench.py --batch 1024 --steps 1 --warmup 3 --no-cpu-baseline --no-parity --no-graph"
$CMD > gpurun_out/${tag}_hbm_plain.log 2>&1 &&
EADGAN_PROFILE_WINDOW=1 timeout 900 ncu --profile-from-start off --metrics $M --clock-control none --csv \
    --log-file gpurun_out/${tag}_step_kernels.csv $CMD > gpurun_out/${tag}_hbm_ncu.log 2>&1
python tools/bench_layers.py 1024 "" 2 > gpurun_out/${tag}_layers_plain.log 2>&1 &&
timeout 900 ncu --metrics $M --clock-control none --csv -k regex:"tc_conv_kernel|tc_wgrad_kernel|tc_dgradT_kernel" \
    --log-file gpurun_out/${tag}_tensor.csv python tools/bench_layers.py 1024 "" 1 > gpurun_out/${tag}_tensor_ncu.log 2>&1
ls -la gpurun_out | grep ${tag}
