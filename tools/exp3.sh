timeout 900 python -m pytest tests/test_tc_gpu.py tests/test_chain_gpu.py tests/test_step_gpu.py tests/test_graph_gpu.py tests/test_b1024_gpu.py -q -m gpu -x 2>&1 | tail -2
run() { env "$@" python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-parity 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$*', d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'])"; }
run EADGAN_X=0
run EADGAN_DEFER_D=1
run EADGAN_PACK_PREFETCH=0
run EADGAN_PACK_PREFETCH=0 EADGAN_DEFER_D=1
run EADGAN_X=0
