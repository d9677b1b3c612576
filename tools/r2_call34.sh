#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_hamming.py -m gpu -x -q 2>&1 | tail -3
python - <<'PY'
import sys, json
sys.path.insert(0,'.')
import bench
peaks,_ = bench.measured_peaks()
print(json.dumps(bench.leg_hamming(0, peaks))[:900])
PY
