# 2-GPU round-end checks: guard-band tests (1 GPU), data-parallel tests, N=2 bench line with dp_parity, N=2 timeline
timeout 600 python -m pytest tests/test_guard_gpu.py -q -m gpu -x 2>&1 | tail -3
timeout 900 python -m pytest tests/test_dp_gpu.py -q -m gpu -x 2>&1 | tail -2
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02w_bench_2gpu.jsonl 2> gpurun_out/r02w_bench_2gpu.err
python -c "
import json; d=json.loads(open('gpurun_out/r02w_bench_2gpu.jsonl').read().strip().splitlines()[-1]); print('N=2', d['value'], d['ms_per_step'], d['e2e']['value'], d.get('dp_parity'))"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 tools/timeline.py --out gpurun_out/r02w_timeline_2gpu.txt 2>&1 | grep -A14 "collectives"
