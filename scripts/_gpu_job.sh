python -m pytest tests/test_reference_gpu.py tests/test_rollout_gpu.py -x -q > gpurun_out/r2o_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/r2o_tests.log
tail -30 gpurun_out/r2o_tests.log
ncu --set full --clock-control none --import-source on -k regex:sa_step_kernel -s 20 -c 1 -f -o gpurun_out/prof_sa_r2n_discrete python scripts/perf_sa.py 4194304 > gpurun_out/r2o_ncu_sa.log 2>&1
tail -5 gpurun_out/r2o_ncu_sa.log
