#!/bin/bash
# round 2, call 21: register cap of the wide 8-lane classes (A/B), smoke()
cd /root/repo
mkdir -p gpurun_out
python __graft_entry__.py smoke 2>&1 | tail -2
for v in default l8w2; do
  if [ $v = default ]; then unset RSA_EXT_LIB; else export RSA_EXT_LIB=/root/repo/variants/librsa_ext_$v.so; fi
  python - <<'PY'
import json, os, sys
sys.path.insert(0, '.')
import bench
r = bench.leg_250bp_indel(0, 0)
print(os.environ.get('RSA_EXT_LIB', 'default'), 'indel leg resident', round(r['resident_gcups']), 'e2e', round(r['e2e_gcups_pageable_host']), r['records_equal'])
PY
  timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra-legs --read-len 250 --pairs 524288 > gpurun_out/r2c21_bench_250_$v.json 2>> gpurun_out/r2c21.err
  python -c "
import json
d=json.load(open('gpurun_out/r2c21_bench_250_$v.json')); print('$v', '250bp value', round(d['value']), 'e2e', round(d['e2e']['value']), 'dp', round(d['roofline']['achieved']))"
done
