python -m pytest tests/test_ma_gpu.py tests/test_reference_gpu.py tests/test_configs_gpu.py -x -q -k "thread or golden or vecenv or ma or extension" > gpurun_out/r3j_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/r3j_tests.log
tail -4 gpurun_out/r3j_tests.log
python scripts/perf_ma.py 65536 131072 262144 2>&1 | tail -5
ncu --set full --clock-control none --import-source on -k regex:ma_tpe_step_kernel -s 40 -c 1 -f -o gpurun_out/prof_ma_tpe_r3j python scripts/perf_ma.py 262144 > gpurun_out/r3j_ncu_ma.log 2>&1
