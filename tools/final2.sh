# round-end evidence on one GPU (small outputs only): serialised kernel list of one eager step with DRAM bytes, tensor-pipe
# list of the layer micro-benchmark, launch list of one graph-replayed step, one source-level full capture of the dominant
# kernel, device timeline of a graph replay, compute-sanitizer memcheck of the smoke step
tag=${1:-r02w}
bash tools/profile_hbm.sh $tag 2>&1 | tail -3
python tools/bench_layers.py 1024 "" 10 > gpurun_out/${tag}_layer_microbench.txt 2>&1
python tools/bench_thin.py 1024 10 >> gpurun_out/${tag}_layer_microbench.txt 2>&1
EADGAN_PROFILE_WINDOW=1 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none \
    --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --batch 1024 --steps 1 --warmup 3 --no-cpu-baseline --no-parity \
    > gpurun_out/${tag}_ncu_launches.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tc_dgradT_kernel -s 2 -c 1 -f -o gpurun_out/${tag}_full_0 \
    python tools/bench_layers.py 1024 "=dgrad+mask" 3 128 > gpurun_out/${tag}_ncu_full_0.log 2>&1
python tools/timeline.py --out gpurun_out/${tag}_timeline.txt > /dev/null 2>&1
export EADGAN_TC_UNITS=8
timeout 420 compute-sanitizer --tool memcheck --log-file gpurun_out/${tag}_memcheck_smoke.log python __graft_entry__.py --smoke > gpurun_out/${tag}_memcheck_smoke.out 2>&1
echo "memcheck smoke rc=$?"; tail -2 gpurun_out/${tag}_memcheck_smoke.log; tail -2 gpurun_out/${tag}_memcheck_smoke.out
timeout 300 compute-sanitizer --tool memcheck --log-file gpurun_out/${tag}_memcheck_tc.log python -m pytest tests/test_tc_gpu.py -q -m gpu -x -k "test_transposed_dgrad or test_channel_major_thin_fprop or test_conv_cta_pairs" > gpurun_out/${tag}_memcheck_tc.out 2>&1
echo "memcheck tests rc=$?"; tail -2 gpurun_out/${tag}_memcheck_tc.out; tail -2 gpurun_out/${tag}_memcheck_tc.log
ls -la gpurun_out | grep $tag
