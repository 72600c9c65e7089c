# NOTE: compute-sanitizer is closed on the GPU pool used this round (gpurun answers "compute-sanitizer is closed on this pool");
# kept for pools where it is available.  The write-bounds check that replaces it is tests/test_guard_gpu.py.
# compute-sanitizer memcheck over the smoke step and the small-geometry kernel parity tests (bounded: the tool slows
# kernels by 10-100x).  Writes gpurun_out/r02z_memcheck_*.log; "ERROR SUMMARY: 0 errors" is the pass line.
export EADGAN_TC_UNITS=8     # small persistent grids: same code paths, far fewer instrumented threads
timeout 500 compute-sanitizer --tool memcheck --log-file gpurun_out/r02z_memcheck_smoke.log python __graft_entry__.py --smoke > gpurun_out/r02z_smoke.out 2>&1
echo "smoke rc=$?"; tail -2 gpurun_out/r02z_memcheck_smoke.log
timeout 600 compute-sanitizer --tool memcheck --log-file gpurun_out/r02z_memcheck_tc.log python -m pytest tests/test_tc_gpu.py tests/test_thin_gpu.py tests/test_ops_gpu.py -q -m gpu -x -k "not 1024 and not large" > gpurun_out/r02z_tc.out 2>&1
echo "tests rc=$?"; tail -2 gpurun_out/r02z_tc.out; tail -2 gpurun_out/r02z_memcheck_tc.log
