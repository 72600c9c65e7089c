# 8-GPU round-end numbers: weak scaling (1024 per GPU, with dp_parity) and the global-1024 (128 per GPU) configuration
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 500 $TR --master-port 29551 bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02w_bench_8gpu.jsonl 2> gpurun_out/r02w_bench_8gpu.err
timeout 300 $TR --master-port 29552 bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu-baseline --no-parity --global-batch 1024 > gpurun_out/r02w_bench_8gpu_global1024.jsonl 2>> gpurun_out/r02w_bench_8gpu.err
python - <<'PY'
import json
for f in ("gpurun_out/r02w_bench_8gpu.jsonl", "gpurun_out/r02w_bench_8gpu_global1024.jsonl"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        dp = d.get("dp_parity") or {}
        print(f, d["value"], d["ms_per_step"], d["e2e"]["value"], d["config"].get("global_batch"), {k: dp.get(k) for k in ("pass", "loss_max_abs_err")})
    except Exception as e:
        print(f, "failed", e)
PY
tail -3 gpurun_out/r02w_bench_8gpu.err | cut -c1-300
