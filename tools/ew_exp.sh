# channel-major kernel: epilogue warps (8 / 16) and compile-time specialisation (on / off)
timeout 400 python -m pytest tests/test_tc_gpu.py tests/test_thin_gpu.py -q -m gpu -x 2>&1 | tail -2
EADGAN_TC_EW=8 timeout 300 python -m pytest tests/test_tc_gpu.py tests/test_thin_gpu.py -q -m gpu -x 2>&1 | tail -1
for ew in 8 16; do for spec in 0 1; do
  echo "== EW=$ew SPEC=$spec"
  EADGAN_TC_EW=$ew EADGAN_TC_SPEC=$spec python tools/bench_thin.py 1024 10 2>&1 | grep fprop
  EADGAN_TC_EW=$ew EADGAN_TC_SPEC=$spec python tools/bench_layers.py 1024 "dgrad" 10 128 2>&1 | grep -v "dgrad+bias"
done; done
for ew in 8 16; do
EADGAN_TC_EW=$ew python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-parity 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('EW=$ew step', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['kernel'], d['roofline']['frac'])"
done
