#!/bin/bash
set -x
cd /root/repo
timeout 900 python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r2c9_pytest.txt 2>&1
tail -5 gpurun_out/r2c9_pytest.txt
./tools/dpx_microbench > gpurun_out/r2c9_dpx_microbench.jsonl 2> gpurun_out/r2c9_dpx.err
grep -E "recipe|HSET2|LOP3\"|VIADDMNMX.S16x2\"" gpurun_out/r2c9_dpx_microbench.jsonl
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2c9_bench.json 2> gpurun_out/r2c9_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2c9_bench.json'))
print("value", round(d['value']), "e2e", round(d['e2e']['value']), "leg250", d['detail']['leg_250bp_5pct_indel'])
PY
