#!/bin/bash
# round 2, call 51: the final default bench line and reference arm on the final tree
cd /root/repo
mkdir -p gpurun_out
T0=$(date +%s); timeout 1200 python bench.py > gpurun_out/r2c51_bench.json 2> gpurun_out/r2c51_bench.err; echo "default bench.py run: $(( $(date +%s) - T0 )) s"
timeout 900 python bench.py --impl reference --steps 20 > gpurun_out/r2c51_bench_ref.json 2>/dev/null
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2c51_bench.json'))
print("value", round(d['value']), "e2e", round(d['e2e']['value']), "dp", round(d['roofline']['achieved']), round(d['roofline']['frac'],3), "steps", d['steps'], d['clocks'])
print('hamming', d['detail']['hamming_shortcut']['pairs_per_s_e2e'], 'sam', d['detail']['sam_format']['records_per_s_e2e'])
print('pipeline', json.dumps(d['detail']['pipeline'])[:1200])
r=json.load(open('gpurun_out/r2c51_bench_ref.json')); print('ref', r['value'], r['steps'], r['ms_per_step'])
PY
