python -m pytest tests/test_ma_gpu.py tests/test_reference_gpu.py -x -q -k "thread or golden or vecenv" > gpurun_out/r2t_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/r2t_tests.log
tail -25 gpurun_out/r2t_tests.log
HW_MA_KERNEL=thread python scripts/perf_ma.py 16384 262144 2>&1 | tail -4
python scripts/perf_ma.py 16384 262144 2>&1 | tail -4
