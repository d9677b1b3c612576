#!/bin/bash
# round 2, call 24: A/B of the fact merges (HFMA2 vs IMAD): bare recipe and kernel
cd /root/repo
mkdir -p gpurun_out
for b in tools/dpx_microbench variants/dpx_microbench_imad; do
  timeout 300 ./$b --quick 2>&1 | grep "SW cell recipe" | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('$b', d['warps_per_sm'], round(d['chip_gcups']))"
done
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra-legs"
for v in default imad default imad; do
  if [ $v = default ]; then unset RSA_EXT_LIB; else export RSA_EXT_LIB=/root/repo/variants/librsa_ext_$v.so; fi
  timeout 600 $B > gpurun_out/r2c24_bench_$v.json 2>> gpurun_out/r2c24.err
  python -c "
import json
d=json.load(open('gpurun_out/r2c24_bench_$v.json')); print('$v', 'value', round(d['value']), 'e2e', round(d['e2e']['value']), 'dp', round(d['roofline']['achieved']), d['detail']['resident_equals_e2e_records'], d['detail']['records_sane'])"
done
