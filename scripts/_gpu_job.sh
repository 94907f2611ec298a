python - <<'PY'
import sys
sys.path.insert(0, 'scripts'); sys.path.insert(0, '.')
import perf_ma
for A in (1, 2, 3, 4):
    for n in (4096, 16384):
        for k in ('thread', 'group'):
            print(k, end=' '); perf_ma.probe(n, A=A, kernel=k)
PY
