#!/bin/bash
# round 2, GPU call 4: seeding kernels -- parity vs the reference's seeding path, first throughput probe
set -x
cd /root/repo
timeout 1200 python -m pytest tests/test_gpu_seed.py -m gpu -x -q > gpurun_out/r2c4_pytest.txt 2>&1
tail -25 gpurun_out/r2c4_pytest.txt
timeout 600 python tools/seed_probe.py 1000000 5000000 > gpurun_out/r2c4_seed_probe.json 2> gpurun_out/r2c4_seed_probe.err
cat gpurun_out/r2c4_seed_probe.json; tail -3 gpurun_out/r2c4_seed_probe.err
timeout 300 python bench.py --steps 10 --warmup 3 --no-extra-legs --no-cpu-baseline > gpurun_out/r2c4_bench.json 2> gpurun_out/r2c4_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2c4_bench.json'))
print("value", round(d['value']), "e2e", round(d['e2e']['value']), "ratio", round(d['e2e']['value']/d['value'],3), "plan_ms", round(d['e2e']['host_plan_ms_per_step'],2))
PY
