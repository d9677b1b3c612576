#!/bin/bash
# round 2, GPU call 6: pipeline with GPU seeding: SAM identity + reads/s
set -x
cd /root/repo
timeout 1500 python -m pytest tests/test_gpu_sam.py -m gpu -x -q > gpurun_out/r2c6_pytest_sam.txt 2>&1
tail -8 gpurun_out/r2c6_pytest_sam.txt
timeout 900 python tools/e2e_reads_bench.py --ref-len 50000000 --reads 3000000 --threads 16 --repeat 2 \
   --binaries rabbitsalign_gasalgpu,rabbitsalign_b200_big,rabbitsalign_b200_win,rabbitsalign_b200_gpuseed \
   > gpurun_out/r2c6_e2e_se_3m.json 2> gpurun_out/r2c6_e2e_se_3m.err
cat gpurun_out/r2c6_e2e_se_3m.json
tail -3 gpurun_out/r2c6_e2e_se_3m.err
