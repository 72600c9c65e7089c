# round-end validation on one GPU: full GPU test suite, smoke, default bench line (parity + cpu baseline), reference arm
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 600 python __graft_entry__.py --smoke 2>&1 | tail -4
timeout 900 python bench.py > gpurun_out/r02w_bench_b1024.jsonl 2> gpurun_out/r02w_bench.err; tail -c 3000 gpurun_out/r02w_bench_b1024.jsonl
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02w_bench_reference.jsonl 2>> gpurun_out/r02w_bench.err; cat gpurun_out/r02w_bench_reference.jsonl
