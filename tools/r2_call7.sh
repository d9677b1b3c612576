#!/bin/bash
set -x
cd /root/repo
timeout 1200 python -m pytest tests/test_gpu_seed.py -m gpu -x -q > gpurun_out/r2c7_pytest.txt 2>&1
tail -12 gpurun_out/r2c7_pytest.txt
timeout 300 python tools/seed_probe.py 1000000 5000000 0 > gpurun_out/r2c7_seed_probe_norepeat.json 2> gpurun_out/r2c7_seed_probe.err
cat gpurun_out/r2c7_seed_probe_norepeat.json
timeout 300 python tools/seed_probe.py 1000000 5000000 8 > gpurun_out/r2c7_seed_probe_repeat.json 2>> gpurun_out/r2c7_seed_probe.err
cat gpurun_out/r2c7_seed_probe_repeat.json
timeout 300 python tools/seed_probe.py 10000 5000000 8 > gpurun_out/r2c7_seed_probe_10k.json 2>> gpurun_out/r2c7_seed_probe.err
cat gpurun_out/r2c7_seed_probe_10k.json
tail -3 gpurun_out/r2c7_seed_probe.err
