#!/bin/bash
# round 2, call 31: 8x8-unit direction tiles + larger handle pool -- parity, SAM identity, bench, traceback traffic,
# 250-bp indel-rich pipeline (3 runs per build)
cd /root/repo
mkdir -p gpurun_out
timeout 2400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py tests/test_gpu_alninfo.py tests/test_gpu_reference_gpu.py tests/test_gpu_sam.py -m gpu -x -q > gpurun_out/r2c31_pytest.txt 2>&1
tail -4 gpurun_out/r2c31_pytest.txt
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra-legs > gpurun_out/r2c31_bench.json 2> gpurun_out/r2c31.err
python -c "
import json
d=json.load(open('gpurun_out/r2c31_bench.json')); print('value', round(d['value']), 'e2e', round(d['e2e']['value']), 'dp', round(d['roofline']['achieved']), 'tb_ms', round(d['roofline']['tb_ms'],2), d['detail']['resident_equals_e2e_records'], d['detail']['records_sane'])"
python - <<'PY'
import sys
sys.path.insert(0,'.')
import bench
r = bench.leg_250bp_indel(0, 0)
print('indel leg resident', round(r['resident_gcups']), 'e2e', round(r['e2e_gcups_pageable_host']), r['records_equal'])
PY
S="python bench.py --pairs 262144 --steps 2 --warmup 3 --no-cpu-baseline --no-extra-legs"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2c31_launches.csv $S > gpurun_out/r2c31_ncu1.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2c31_launches.csv')) if len(r)>10]
hdr=rows[0]; k=hdr.index('Kernel Name'); m=hdr.index('Metric Name'); v=hdr.index('Metric Value'); i=hdr.index('ID')
cur={}
for r in rows[1:]:
    cur.setdefault(r[i],{'name':r[k].split('(')[0][:30]})[r[m]]=float(r[v].replace(',',''))
for pat in ('tb_groups','fast_dp'):
    big=sorted([d for d in cur.values() if pat in d['name']], key=lambda d:-d['gpu__time_duration.sum'])
    for d in big[:2]: print(d)
PY
timeout 1200 python tools/e2e_reads_bench.py --ref-len 50000000 --reads 1000000 --threads $(nproc) --repeat 3 --read-len 250 --sub 0.02 --indel 0.02 --max-indel 4 \
  --binaries rabbitsalign_gasalgpu,rabbitsalign_b200_big,rabbitsalign_b200_gpuseed > gpurun_out/r2c31_e2e_se_250bp.json 2> gpurun_out/r2c31_e2e.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2c31_e2e_se_250bp.json'))
for k,v in d.items():
    if isinstance(v,dict): print(' ', k, v.get('wall_s_runs'), v.get('mapping_s'), v.get('sam_md5'), v.get('error'))
PY
