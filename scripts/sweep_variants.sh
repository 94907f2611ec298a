#!/bin/bash
# run the smoke parity check and scripts/perf_sa.py once per prebuilt library variant in build/variants/ (tuning experiments)
keep=$(mktemp); cp gym_highway_b200/libhighway_b200.so $keep
for f in build/variants/*.so; do
  cp $f gym_highway_b200/libhighway_b200.so
  echo "== $f"
  if [ -z "$NO_SMOKE" ]; then python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1; fi
  python scripts/perf_sa.py ${1:-4194304} 2>&1 | tail -2
done
cp $keep gym_highway_b200/libhighway_b200.so
