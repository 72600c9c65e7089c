#!/bin/bash
# GPU-box profiling pass (run under gpurun): launch list of one graph-replayed bench step + ncu --set full of
# the layer kernels named on the command line.  usage: tools/profile_round.sh <tag> ["<filter> <c>" ...]
set -u
tag=${1:-rXX}; shift
mkdir -p gpurun_out
python bench.py --batch 1024 --steps 1 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/${tag}_plain.log 2>&1 &&
EADGAN_PROFILE_WINDOW=1 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none \
    --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --batch 1024 --steps 1 --warmup 3 --no-cpu-baseline --no-parity \
    > gpurun_out/${tag}_ncu_launches.log 2>&1
i=0
for spec in "$@"; do
  set -- $spec
  flt=$1; c=$2; kern=${3:-tc_conv_kernel}
  python tools/bench_layers.py 1024 "$flt" 3 $c > gpurun_out/${tag}_layer_${i}.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:$kern -s 2 -c 2 -f -o gpurun_out/${tag}_full_${i} \
      python tools/bench_layers.py 1024 "$flt" 3 $c > gpurun_out/${tag}_ncu_full_${i}.log 2>&1
  i=$((i+1))
done
ls -la gpurun_out | tail -20
