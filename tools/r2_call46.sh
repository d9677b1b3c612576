#!/bin/bash
# round 2, call 46: per-array bounce decision in the Hamming entry points: parity + the bench leg
cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_hamming.py tests/test_gpu_sam_format.py -m gpu -q > gpurun_out/r2c46_pytest.txt 2>&1; tail -2 gpurun_out/r2c46_pytest.txt
python - <<'PY'
import sys, json
sys.path.insert(0, '.')
import bench
print(json.dumps(bench.leg_hamming(0, {'hbm_gbs': 6471.1}) if hasattr(bench, 'leg_hamming') else None)[:900])
PY
