#!/bin/bash
# round 2, call 33: Hamming / SAM tests after the timing additions, the bench line with all legs
cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_hamming.py tests/test_gpu_sam_format.py -m gpu -x -q 2>&1 | tail -2
python __graft_entry__.py smoke 2>&1 | tail -1
timeout 1200 python bench.py --steps 20 --warmup 3 > gpurun_out/r2c33_bench.json 2> gpurun_out/r2c33_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2c33_bench.json'))
print("value", round(d['value']), "e2e", round(d['e2e']['value']), "dp", round(d['roofline']['achieved']), round(d['roofline']['frac'],3), 'cpu', d['cpu_baseline']['value'])
for k in ('leg_250bp_5pct_indel','hamming_shortcut','sam_format','pipeline'):
    print(k, json.dumps(d['detail'].get(k))[:1000])
PY
