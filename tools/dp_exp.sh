# experiment (2 GPUs): SMs reserved for NCCL while gradient all-reduces are in flight x NCCL CTA cap
for cfg in "0 0" "8 0" "8 8" "16 16" "4 4"; do set -- $cfg
  export EADGAN_DP_RESERVE_SMS=$1; if [ "$2" != "0" ]; then export NCCL_MAX_CTAS=$2; else unset NCCL_MAX_CTAS; fi
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 --no-parity 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('reserve $1 nccl_max_ctas $2:', round(d['value']), 'img/s', round(d['ms_per_step'],3), 'ms')"
done
