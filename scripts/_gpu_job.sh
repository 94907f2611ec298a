keep=$(mktemp); cp gym_highway_b200/libhighway_b200.so $keep
for f in build/variants/*.so; do
  cp $f gym_highway_b200/libhighway_b200.so
  echo "== $f"; HW_MA_KERNEL=thread python scripts/perf_ma.py 262144 2>&1 | tail -2; python scripts/perf_ma.py 262144 2>&1 | tail -2; python scripts/perf_sa.py 4194304 2>&1 | tail -2
done
cp $keep gym_highway_b200/libhighway_b200.so
