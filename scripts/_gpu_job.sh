ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r3g_collector_launches.csv python scripts/perf_collector.py > /dev/null 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/r3g_collector_launches.csv')))
hdr=None; seq=[]
for r in rows:
    if 'Kernel Name' in r: hdr=r; continue
    if hdr and len(r)==len(hdr):
        d=dict(zip(hdr,r)); seq.append((d['Kernel Name'][:70], float(d['Metric Value'].replace(',',''))/1e3))
# last step's launches
for k,v in seq[-16:]: print('%8.1f us  %s' % (v,k))
PY
