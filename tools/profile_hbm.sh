#!/bin/bash
# GPU-box profiling pass for the HBM-bound kernels (run under gpurun): `ncu --set full` of every BatchNorm /
# spectral-norm / thin-image / Adam / packing / reduction kernel of ONE eager CelebA step at B=1024.
# usage: tools/profile_hbm.sh <tag>     ->  gpurun_out/<tag>_hbm.ncu-rep ; summarise with tools/ncu_hbm_summarize.py
set -u
tag=${1:-rXX}
mkdir -p gpurun_out
CMD="python bench.py --batch 1024 --steps 1 --warmup 3 --no-cpu-baseline --no-parity --no-graph"
$CMD > gpurun_out/${tag}_hbm_plain.log 2>&1 &&
EADGAN_PROFILE_WINDOW=1 ncu --profile-from-start off --set full --clock-control none \
    -k regex:"bn_|sn_|adam_kernel|thin_|wgrad_reduce|pack_w|dense_pack|copy4|philox|stn_fwd|act_" -c 260 -f \
    -o gpurun_out/${tag}_hbm $CMD > gpurun_out/${tag}_hbm_ncu.log 2>&1
ls -la gpurun_out | grep ${tag}_hbm
