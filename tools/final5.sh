# the other BASELINE configurations on one GPU (with their parity blocks), and the BatchNorm grid knob on them
for cfg in dsprites colored; do
  for bps in 2 8; do
    EADGAN_BN_BLOCKS_PER_SM=$bps timeout 300 python bench.py --config $cfg --steps 30 --warmup 5 --no-cpu-baseline $([ $bps = 8 ] && echo --no-parity) > gpurun_out/r02w_bench_${cfg}_bn$bps.jsonl 2>> gpurun_out/r02w_bench_cfg.err
    python -c "
import json; d=json.loads(open('gpurun_out/r02w_bench_${cfg}_bn$bps.jsonl').read().strip().splitlines()[-1]); print('$cfg bn_blocks_per_sm=$bps', d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'], (d.get('parity') or {}).get('pass'))"
  done
done
