#!/bin/bash
# round 2, call 17: packed reference planes + vectorised window staging -- parity and bench
cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py tests/test_gpu_hamming.py -m gpu -x -q > gpurun_out/r2c17_pytest.txt 2>&1
tail -12 gpurun_out/r2c17_pytest.txt
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra-legs > gpurun_out/r2c17_bench.json 2> gpurun_out/r2c17.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2c17_bench.json'))
print("value", round(d['value']), "e2e", round(d['e2e']['value']), "dp", round(d['roofline']['achieved']), d['detail'].get('e2e_windows_in_resident_reference'))
PY
