python -m pytest tests/test_configs_gpu.py tests/test_reference_gpu.py -x -q > gpurun_out/r3e_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/r3e_tests.log
tail -6 gpurun_out/r3e_tests.log
python bench.py > gpurun_out/r3e_bench.json 2> gpurun_out/r3e_bench.err; echo "bench rc $?"; tail -3 gpurun_out/r3e_bench.err
