EADGAN_TC_DGRADT=1 python tools/bench_layers.py 1024 "=dgrad" 3 128 > gpurun_out/r02p_plain.log 2>&1 &&
EADGAN_TC_DGRADT=1 ncu --set full --clock-control none --import-source on -k regex:tc_dgradT_kernel -s 2 -c 1 -f -o gpurun_out/r02p_dgradT python tools/bench_layers.py 1024 "=dgrad" 3 128 > gpurun_out/r02p_ncu.log 2>&1
tail -3 gpurun_out/r02p_ncu.log
