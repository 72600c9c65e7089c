#!/bin/bash
# ncu detail sections of selected non-GEMM kernels inside ONE eager CelebA step at B = 1024 (text log, small)
# usage: tools/prof_sel.sh <tag> "<kernel regex>" [count]
set -u
tag=${1:-rXX}; rx=${2:-bn_bwd_reduce}; cnt=${3:-8}
mkdir -p gpurun_out
CMD="python bench.py --batch 1024 --steps 1 --warmup 3 --no-cpu-baseline --no-parity --no-graph"
EADGAN_PROFILE_WINDOW=1 timeout 900 ncu --profile-from-start off --clock-control none -k regex:"$rx" -c $cnt \
  --section SpeedOfLight --section MemoryWorkloadAnalysis --section Occupancy --section WarpStateStats --section LaunchStats --section SchedulerStats \
  --log-file gpurun_out/${tag}_sel.txt $CMD > gpurun_out/${tag}_sel_ncu.log 2>&1
grep -c "^  [a-z_A-Z:<>0-9, ]*(" gpurun_out/${tag}_sel.txt; ls -la gpurun_out/${tag}_sel.txt
