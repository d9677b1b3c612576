#!/bin/bash
# round 2, call 28: merged (flex) column classes -- parity, indel-leg probe, bench
cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py tests/test_gpu_alninfo.py -m gpu -x -q > gpurun_out/r2c28_pytest.txt 2>&1
tail -3 gpurun_out/r2c28_pytest.txt
python tools/indel_leg_probe.py
python - <<'PY'
import sys
sys.path.insert(0,'.')
import bench
r = bench.leg_250bp_indel(0, 0)
print('indel leg resident', round(r['resident_gcups']), 'e2e', round(r['e2e_gcups_pageable_host']), r['records_equal'])
PY
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra-legs > gpurun_out/r2c28_bench.json 2> gpurun_out/r2c28.err
python -c "
import json
d=json.load(open('gpurun_out/r2c28_bench.json')); print('value', round(d['value']), 'e2e', round(d['e2e']['value']), 'dp', round(d['roofline']['achieved']), d['detail']['resident_equals_e2e_records'], d['detail']['records_sane'], d['gpu_launches'])"
