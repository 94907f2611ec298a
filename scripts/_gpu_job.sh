python -m pytest tests -m gpu -x -q > gpurun_out/r3l_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/r3l_tests.log
tail -5 gpurun_out/r3l_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py > gpurun_out/r3l_bench.json 2> gpurun_out/r3l_bench.err; echo "bench rc $?"; tail -3 gpurun_out/r3l_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r3l_ref.json 2>&1; tail -c 600 gpurun_out/r3l_ref.json
