for k in thread group; do echo "== $k"; HW_MA_KERNEL=$k python scripts/perf_ma.py 16384 24576 32768 49152 65536 2>&1 | head -5; done
