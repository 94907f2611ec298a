keep=$(mktemp); cp gym_highway_b200/libhighway_b200.so $keep
for f in build/variants/*.so; do
  cp $f gym_highway_b200/libhighway_b200.so
  echo "== $f"; python scripts/perf_sa.py 16384 65536 262144 2>&1 | head -3
done
cp $keep gym_highway_b200/libhighway_b200.so
