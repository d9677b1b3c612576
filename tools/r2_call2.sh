#!/bin/bash
# round 2, GPU call 2: device-side planner parity + bench with the new roofline block / legs
set -x
cd /root/repo
timeout 900 python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2c2_pytest.txt
cat gpurun_out/r2c2_pytest.txt
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2c2_bench.json 2> gpurun_out/r2c2_bench.err
tail -c 3000 gpurun_out/r2c2_bench.json
tail -5 gpurun_out/r2c2_bench.err
RSA_EXT_TRACE=1 timeout 120 python tools/e2e_timeline.py > gpurun_out/r2c2_timeline.txt 2>&1
tail -40 gpurun_out/r2c2_timeline.txt
