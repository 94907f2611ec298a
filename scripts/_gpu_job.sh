python -m pytest tests -m gpu -x -q > gpurun_out/r3a_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/r3a_tests.log
tail -5 gpurun_out/r3a_tests.log
python bench.py > gpurun_out/r3a_bench.json 2> gpurun_out/r3a_bench.err; echo "bench rc $?"; tail -5 gpurun_out/r3a_bench.err
HW_MA_KERNEL=thread python - <<'PY'
import os, sys
sys.path.insert(0, 'scripts'); sys.path.insert(0, '.')
import perf_ma
for k in ('thread', 'group'):
    for A in (3, 1):
        perf_ma.probe(262144, A=A, kernel=k)
PY
