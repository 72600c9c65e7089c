timeout 600 python -m pytest tests/test_tc_gpu.py tests/test_thin_gpu.py tests/test_ops_gpu.py tests/test_chain_gpu.py tests/test_step_gpu.py -q -m gpu -x 2>&1 | tail -2
python tools/bench_thin.py 1024 10 2>&1 | grep "fprop\|dgrad"
for bps in 8 4 2; do
EADGAN_BN_BLOCKS_PER_SM=$bps python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-parity 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bn blocks/SM $bps step', d['value'], d['ms_per_step'], d['e2e']['value'])"
done
