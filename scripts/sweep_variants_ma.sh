#!/bin/bash
# run the multi-agent tests and scripts/perf_ma.py once per prebuilt library variant in build/variants/ (tuning experiments)
keep=$(mktemp); cp gym_highway_b200/libhighway_b200.so $keep
for f in build/variants/*.so; do
  cp $f gym_highway_b200/libhighway_b200.so
  echo "== $f"; python -m pytest tests/test_ma_gpu.py -m gpu -x -q 2>&1 | tail -1; python scripts/perf_ma.py ${1:-262144} 2>&1 | tail -2
done
cp $keep gym_highway_b200/libhighway_b200.so
