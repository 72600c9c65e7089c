timeout 100 python -m pytest tests/test_ops_gpu.py tests/test_dsprites_gpu.py tests/test_colored_gpu.py tests/test_mnist_gpu.py tests/test_pxy_gpu.py tests/test_step_gpu.py tests/test_aux_gpu.py -q -m gpu -x 2>&1 | tail -2
timeout 40 python __graft_entry__.py --smoke 2>&1 | tail -2
timeout 40 python bench.py --config dsprites --steps 30 --warmup 5 --no-cpu-baseline --no-parity 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('dsprites', d['value'], d['ms_per_step'])"
