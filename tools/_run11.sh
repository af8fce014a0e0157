mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02_launches_states.csv python tools/states_profile.py > gpurun_out/states_profile.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r02_launches_states.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
import collections
agg=collections.defaultdict(list)
for r in rows[1:]:
    try: agg[r[ki][:60]].append(float(r[vi].replace(',','')))
    except ValueError: pass
for k,v in agg.items(): print(k, len(v), 'avg us', sum(v)/len(v)/1000 if max(v)>1000 else sum(v)/len(v))
PY
python tools/states_profile.py
STATES_BATCH=65536 STATES_STEPS=5 python tools/states_profile.py
