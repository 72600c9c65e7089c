bash tools/sweep.sh r02s 1 > gpurun_out/r02s_sweep_n1.txt 2>&1
cat gpurun_out/r02s_sweep_n1.txt | tail -12
bash tools/profile_hbm.sh r02s > gpurun_out/r02s_hbm.log 2>&1; tail -2 gpurun_out/r02s_hbm.log
bash tools/profile_round.sh r02s "=dgrad 128 tc_dgradT_kernel" "=wgrad 128 tc_wgrad_kernel" "=wgrad 256 tc_wgrad_kernel" "=fprop 128" "=fprop 512" "=dgrad 256" "=dgrad 512" > gpurun_out/r02s_profile.log 2>&1; tail -3 gpurun_out/r02s_profile.log
