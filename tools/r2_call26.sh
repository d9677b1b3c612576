#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
python tools/indel_leg_probe.py
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2c26_indel_launches.csv python tools/indel_leg_probe.py > /dev/null 2>&1
python - <<'PY'
import csv,collections
rows=[r for r in csv.reader(open('gpurun_out/r2c26_indel_launches.csv')) if len(r)>10]
hdr=rows[0]; k=hdr.index('Kernel Name'); v=hdr.index('Metric Value'); g=hdr.index('Grid Size')
agg=collections.defaultdict(list)
for r in rows[1:]:
    agg[(r[k].split('(')[0][:40], r[g])].append(float(r[v].replace(',','')))
for key,vals in sorted(agg.items(), key=lambda x:-sum(x[1])): print(key, len(vals), round(sum(vals)/len(vals)/1e3,1), 'us')
PY
