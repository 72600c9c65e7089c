# experiment (N GPUs): gradient bucket size (one all-reduce per phase vs 32 MB buckets), SM reservation
N=${1:-8}
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 --no-parity --no-cpu-baseline "$@" 2>/dev/null | python -c "
import json,sys,os
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('N=$N bucket_mb', os.environ.get('EADGAN_DP_BUCKET_MB','32'), 'reserve', os.environ.get('EADGAN_DP_RESERVE_SMS','0'), 'args', ' '.join(sys.argv[1:]), ':', round(d['value']), 'img/s', round(d['ms_per_step'],3), 'ms')" "$@"; }
EADGAN_DP_BUCKET_MB=256 run
EADGAN_DP_BUCKET_MB=256 run --global-batch 1024
EADGAN_DP_BUCKET_MB=256 EADGAN_DP_RESERVE_SMS=16 NCCL_MAX_CTAS=16 run
