#!/bin/bash
# BASELINE configs beyond the default bench line (run under gpurun [--gpus N]): one bench.py JSON line per case,
# appended to gpurun_out/<tag>_sweep_n<N>.jsonl.   usage: tools/sweep.sh <tag> <N> [quick]
set -u
tag=$1; N=${2:-1}; quick=${3:-}
out=gpurun_out/${tag}_sweep_n${N}.jsonl
: > $out
run() {  # args: extra bench flags ; env assignments may precede through "env"
  if [ "$N" = "1" ]; then
    timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline "$@" >> $out 2>> gpurun_out/${tag}_sweep_n${N}.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29520 \
        bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline "$@" >> $out 2>> gpurun_out/${tag}_sweep_n${N}.err
  fi
  echo "rc=$? $*" >> gpurun_out/${tag}_sweep_n${N}.err
}
if [ "$quick" = "scale" ]; then
  # multi-GPU measurement set (GPU-minutes are charged N-fold): the default line with dp_parity, GLOBAL batch 1024
  # (= 128 per GPU at N = 8: the strong-scaling point), colored (global batch 512) and the two ablations that say
  # which exchange costs what
  run
  run --global-batch 1024 --no-parity
  run --config colored --no-parity
  EADGAN_DP_ABLATE=grads run --no-parity
  EADGAN_DP_ABLATE=syncbn run --no-parity
elif [ "$quick" = "colored" ]; then
  run --config colored --no-parity
else
# configs[3]: CelebA 1024 per GPU (weak), with parity (N=1) / dp_parity (N>1)
run
# configs[4]: per-GPU batch sweep (weak scaling)
for b in 128 512 4096; do run --batch $b --no-parity; done
if [ "$quick" != "quick" ]; then run --batch 8192 --no-parity; fi
# the strong-scaling point: GLOBAL batch 1024 (128 per GPU at N = 8)
if [ "$N" != "1" ]; then run --global-batch 1024; fi
# configs[1] dSprites batch 256 per GPU; configs[2] colored, GLOBAL batch 512
run --config dsprites
run --config colored
fi
python - <<PY
import json
for ln in open("$out"):
    try: d = json.loads(ln)
    except Exception: continue
    c = d["config"]
    print(f"{c['workload'][:34]:34s} B/gpu {c['batch_per_gpu']:5d} global {c['global_batch']:6d} N {d['n_gpus']} {d['scaling']:6s} "
          f"{d['value']:10.0f} img/s  {d['ms_per_step']:8.3f} ms  e2e {d['e2e']['value']:10.0f}  "
          f"parity {d.get('parity', {}).get('pass')} dp_parity {d.get('dp_parity', {}).get('pass')}")
PY
