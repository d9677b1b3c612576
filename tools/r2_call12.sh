#!/bin/bash
# round 2, call 12: clamp-fact recipe -- microbenchmark, parity, bench
set -x
cd /root/repo
mkdir -p gpurun_out
timeout 300 ./tools/dpx_microbench > gpurun_out/r2c12_dpx_microbench.jsonl 2> gpurun_out/r2c12_dpx.err
tail -3 gpurun_out/r2c12_dpx_microbench.jsonl
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py tests/test_gpu_alninfo.py tests/test_gpu_reference_gpu.py -m gpu -x -q > gpurun_out/r2c12_pytest.txt 2>&1
tail -5 gpurun_out/r2c12_pytest.txt
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra-legs > gpurun_out/r2c12_bench.json 2> gpurun_out/r2c12_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2c12_bench.json'))
print("value", round(d['value']), "e2e", round(d['e2e']['value']), "roofline", d['roofline']['achieved'], d['roofline']['frac'], "tb_ms", d['roofline'].get('tb_ms'), "dp_ms", d['roofline'].get('dp_ms'))
PY
