set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r2n_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/r2n_tests.log
python bench.py > gpurun_out/r2n_bench.json 2> gpurun_out/r2n_bench.err; echo "bench exit $?" >> gpurun_out/r2n_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2n_ref.json 2> gpurun_out/r2n_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2n_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/r2n_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sa_step_kernel -s 160 -c 1 -f -o gpurun_out/prof_sa_r2n_discrete python scripts/perf_sa.py 4194304 > gpurun_out/r2n_ncu_sa.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:ma_grp_step_kernel -s 40 -c 1 -f -o gpurun_out/prof_ma_r2n python scripts/perf_ma.py 262144 > gpurun_out/r2n_ncu_ma.log 2>&1
tail -3 gpurun_out/r2n_tests.log; cut -c1-400 gpurun_out/r2n_bench.json
