# one GPU call: the GPU suite, both bench arms, an ncu launch list of the bench, --set full captures of the two step kernels, the soak.
# usage: scripts/gpu_round_check.sh TAG   (files land in gpurun_out/TAG_*)
T=${1:-r4}
set -x
python -m pytest tests -m gpu -x -q > gpurun_out/${T}_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/${T}_tests.log
python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench exit $?" >> gpurun_out/${T}_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_ref.json 2> gpurun_out/${T}_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 1 --no-extras > gpurun_out/${T}_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sa_step_kernel -s 20 -c 1 -f -o gpurun_out/prof_sa_${T}_discrete python scripts/perf_sa.py 4194304 > gpurun_out/${T}_ncu_sa.log 2>&1
HW_MA_KERNEL=thread ncu --set full --clock-control none --import-source on -k regex:ma_tpe_step_kernel -s 35 -c 1 -f -o gpurun_out/prof_tpe_${T} python scripts/perf_ma.py 262144 > gpurun_out/${T}_ncu_tpe.log 2>&1
python scripts/long_parity.py gpurun_out/${T}_long_parity.json > gpurun_out/${T}_long_parity.log 2>&1
tail -3 gpurun_out/${T}_tests.log; cut -c1-300 gpurun_out/${T}_bench.json; tail -3 gpurun_out/${T}_long_parity.log
