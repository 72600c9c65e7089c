# thin image-layer kernels: CUDA-event timings, then one source-level ncu capture of the channel-major fprop
python -m pytest tests/test_step_gpu.py tests/test_chain_gpu.py -q -m gpu -x 2>&1 | tail -2
python tools/bench_thin.py 1024 10 2>&1 | tee gpurun_out/r02y_thin.log &&
ncu --set full --clock-control none --import-source on -k regex:tc_dgradT_kernel -s 2 -c 1 -f -o gpurun_out/r02y_thin python tools/bench_thin.py 1024 2 > gpurun_out/r02y_ncu.log 2>&1
tail -3 gpurun_out/r02y_ncu.log
ls -la gpurun_out/*.ncu-rep
