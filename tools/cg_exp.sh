# channel-major kernel with the next tile's mask prefetched into L2
timeout 120 python tools/bench_layers.py 1024 "dgrad" 10 128 2>&1 | grep -E "dgrad  |stats|mask"
timeout 300 python -m pytest tests/test_tc_gpu.py tests/test_b1024_gpu.py tests/test_ops_gpu.py tests/test_dsprites_gpu.py tests/test_colored_gpu.py tests/test_mnist_gpu.py -q -m gpu -x 2>&1 | tail -2
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-parity 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bench', d['value'], d['ms_per_step'], d['roofline']['kernel'], d['roofline']['frac'])"
for c in dsprites colored; do timeout 300 python bench.py --config $c --steps 20 --warmup 5 --no-cpu-baseline --no-parity 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$c', d['value'], d['ms_per_step'], d['gpu_launches']/20)"; done
