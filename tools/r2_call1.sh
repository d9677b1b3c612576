#!/bin/bash
# round 2, GPU call 1: parity after the advisor fixes, new SAM variants, pipeline reads/s with the new builds
set -x
nvidia-smi -L; nproc; free -g | head -2
cd /root/repo
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2c1_pytest.txt
cat gpurun_out/r2c1_pytest.txt
timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/r2c1_bench.json 2> gpurun_out/r2c1_bench.err
tail -c 1500 gpurun_out/r2c1_bench.json
timeout 600 python tools/e2e_reads_bench.py --ref-len 50000000 --reads 3000000 --threads 16 --repeat 2 \
   --binaries rabbitsalign_gasalgpu,rabbitsalign_b200,rabbitsalign_b200_alninfo,rabbitsalign_b200_big,rabbitsalign_b200_win \
   > gpurun_out/r2c1_e2e_se_3m.json 2> gpurun_out/r2c1_e2e_se_3m.err
cat gpurun_out/r2c1_e2e_se_3m.json
