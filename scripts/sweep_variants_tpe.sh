#!/bin/bash
# time scripts/perf_ma.py with the thread-per-env multi-agent kernel once per prebuilt library variant in build/variants/
keep=$(mktemp); cp gym_highway_b200/libhighway_b200.so $keep
for f in build/variants/*.so; do
  cp $f gym_highway_b200/libhighway_b200.so
  echo "== $f"; HW_MA_KERNEL=thread python scripts/perf_ma.py ${1:-262144} 2>&1 | tail -2
done
cp $keep gym_highway_b200/libhighway_b200.so
