# channel-major kernel: c = 128 dgrad (mask loads pipelined) + thin fprop
for t in 1 0; do EADGAN_TC_DGRADT=$t timeout 120 python tools/bench_layers.py 1024 "dgrad" 10 128 2>&1 | grep -E "dgrad  |mask" | sed "s/^/T$t /"; done
timeout 300 python -m pytest tests/test_tc_gpu.py tests/test_thin_gpu.py tests/test_b1024_gpu.py tests/test_chain_gpu.py -q -m gpu -x 2>&1 | tail -2
for t in 1 0; do EADGAN_TC_DGRADT=$t timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-parity --profile-out gpurun_out/r02r_entry_$t.json 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bench channel-major=$t', d['value'], d['ms_per_step'], d['roofline']['kernel'], d['roofline']['frac'])"; done
python -c "
import json
for t in (1,0):
    d=json.load(open('gpurun_out/r02r_entry_%d.json'%t))['entry_points']
    print(t, {k:(round(v['ms'],3),v['calls']) for k,v in d.items() if 'thin' in k or 'c128' in k})"
