#!/bin/bash
# one GPU call: time every build/variants/*.so - the thread-per-env MA kernel and the SA kernel; parity subsets for $PARITY variants
keep=$(mktemp); cp gym_highway_b200/libhighway_b200.so $keep
for f in build/variants/*.so; do
  cp $f gym_highway_b200/libhighway_b200.so
  echo "== $f"
  if [ -z "$NO_MA" ]; then HW_MA_KERNEL=thread python scripts/perf_ma.py ${SIZES:-262144} 2>&1 | tail -${TAILN:-2}; fi
  if [ -n "$SA" ]; then python scripts/perf_sa.py 65536 4194304 2>&1 | tail -3; fi
  for p in $PARITY; do
    if [ "$(basename $f .so)" == "$p" ]; then
      python -m pytest tests/test_ma_gpu.py tests/test_configs_gpu.py -x -q -m gpu -k "thread and (ma5 or 5-False or 5-True or configs2)" 2>&1 | tail -3
    fi
  done
  for p in $PARITY_SA; do
    if [ "$(basename $f .so)" == "$p" ]; then
      python -m pytest tests/test_sa_gpu.py -x -q -m gpu 2>&1 | tail -3
    fi
  done
done
cp $keep gym_highway_b200/libhighway_b200.so
