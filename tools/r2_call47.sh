#!/bin/bash
# round 2, call 47: pinned-vs-pageable caller tests, smoke(), and the final default bench line (100 timed steps) + reference arm
cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_hamming.py tests/test_gpu_sam_format.py -m gpu -q > gpurun_out/r2c47_pytest.txt 2>&1; tail -3 gpurun_out/r2c47_pytest.txt
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
T0=$(date +%s); timeout 1200 python bench.py > gpurun_out/r2c47_bench.json 2> gpurun_out/r2c47_bench.err; echo "default bench.py run: $(( $(date +%s) - T0 )) s"
T0=$(date +%s); timeout 900 python bench.py --impl reference > gpurun_out/r2c47_bench_ref.json 2>/dev/null; echo "default reference arm: $(( $(date +%s) - T0 )) s"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2c47_bench.json'))
print("value", round(d['value']), "e2e", round(d['e2e']['value']), "dp", round(d['roofline']['achieved']), round(d['roofline']['frac'],3), "steps", d['steps'], d['clocks'])
for k in ('leg_250bp_5pct_indel','e2e_windows_in_resident_reference'): print(k, json.dumps(d['detail'].get(k))[:300])
print('hamming', d['detail']['hamming_shortcut']['pairs_per_s_e2e'], 'sam', d['detail']['sam_format']['records_per_s_e2e'])
print('pipeline', json.dumps(d['detail']['pipeline'])[:700])
r=json.load(open('gpurun_out/r2c47_bench_ref.json')); print('ref', r['value'], r['steps'], r['ms_per_step'])
PY
