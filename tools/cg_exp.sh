# experiment: cta_group::2 vs ::1 on one layer (usage under gpurun: bash tools/cg_exp.sh)
for cg in 2 1; do for f in "=fprop" "=dgrad" "=wgrad"; do EADGAN_TC_CG=$cg timeout 120 python tools/bench_layers.py 1024 "$f" 5 256 2>&1 | tail -1 | sed "s/^/cg$cg /"; done; done
EADGAN_TC_CG=2 timeout 200 python -m pytest tests/test_tc_gpu.py -q -m gpu -x -k "pairs" 2>&1 | tail -2
