#!/bin/bash
# round 2, call 14: A/B builds of the DP kernel (register cap, unroll, selectors in shared memory)
cd /root/repo
mkdir -p gpurun_out
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
for v in w3 w3u2 w2q w3q w3qu2 w4q; do
  RSA_EXT_LIB=/root/repo/variants/librsa_ext_$v.so timeout 600 $B > gpurun_out/r2c14_bench_$v.json 2>> gpurun_out/r2c14.err; show gpurun_out/r2c14_bench_$v.json
done
RSA_EXT_LIB=/root/repo/variants/librsa_ext_w4q.so timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
