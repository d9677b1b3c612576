#!/bin/bash
# round 2, call 15: lane-major direction tiles with 256-bit stores -- parity, bench, traceback traffic
cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py tests/test_gpu_alninfo.py tests/test_gpu_reference_gpu.py -m gpu -x -q > gpurun_out/r2c15_pytest.txt 2>&1
tail -3 gpurun_out/r2c15_pytest.txt
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra-legs"
show() { python - "$1" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print(sys.argv[1], "value", round(d['value']), "e2e", round(d['e2e']['value']), "dp", round(d['roofline']['achieved']), "tb_ms", round(d['roofline'].get('tb_ms',0),2), "eq", d['detail'].get('resident_equals_e2e_records'), d['detail'].get('records_sane'))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
}
timeout 600 $B > gpurun_out/r2c15_bench.json 2> gpurun_out/r2c15.err; show gpurun_out/r2c15_bench.json
timeout 600 $B --read-len 250 --pairs 524288 > gpurun_out/r2c15_bench_250.json 2>> gpurun_out/r2c15.err; show gpurun_out/r2c15_bench_250.json
S="python bench.py --pairs 262144 --steps 2 --warmup 3 --no-cpu-baseline --no-extra-legs"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2c15_launches.csv $S > gpurun_out/r2c15_ncu1.log 2>&1
python - <<'PY'
import csv,collections
rows=[r for r in csv.reader(open('gpurun_out/r2c15_launches.csv')) if len(r)>10]
hdr=rows[0]; k=hdr.index('Kernel Name'); m=hdr.index('Metric Name'); v=hdr.index('Metric Value')
agg=collections.defaultdict(lambda: collections.defaultdict(float)); n=collections.Counter()
for r in rows[1:]:
    name=r[k].split('(')[0][:40]
    agg[name][r[m]]+=float(r[v].replace(',',''))
    if r[m]=='gpu__time_duration.sum': n[name]+=1
for name,d in agg.items(): print(name, n[name], {a:round(b/max(n[name],1),1) for a,b in d.items()})
PY
