for cfg in dsprites colored; do
  timeout 45 python bench.py --config $cfg --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r02w_bench_${cfg}.jsonl 2>/dev/null
  python -c "
import json; d=json.loads(open('gpurun_out/r02w_bench_${cfg}.jsonl').read().strip().splitlines()[-1]); print('$cfg', d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'], (d.get('parity') or {}).get('pass'))"
done
