#!/bin/bash
# round 2, call 19: SAM formatter parity + packed-staging test after the fix
cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sam_format.py tests/test_gpu_round2.py -m gpu -x -q > gpurun_out/r2c19_pytest.txt 2>&1
tail -15 gpurun_out/r2c19_pytest.txt
