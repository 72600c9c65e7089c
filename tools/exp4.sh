timeout 900 python -m pytest tests/test_tc_gpu.py tests/test_chain_gpu.py tests/test_step_gpu.py tests/test_graph_gpu.py tests/test_b1024_gpu.py tests/test_run_gpu.py -q -m gpu -x 2>&1 | tail -2
python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('step', d['value'], d['ms_per_step'], d['e2e']['value'], d['parity']['pass'])"
