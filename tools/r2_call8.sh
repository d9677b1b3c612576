#!/bin/bash
# round 2, GPU call 8: full GPU suite, full bench (all legs), pipeline at configs[1] scale with phase timers
set -x
cd /root/repo
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r2c8_pytest.txt 2>&1
tail -6 gpurun_out/r2c8_pytest.txt
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2c8_bench.json 2> gpurun_out/r2c8_bench.err
tail -c 2500 gpurun_out/r2c8_bench.json; tail -3 gpurun_out/r2c8_bench.err
timeout 120 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2c8_bench_ref.json 2>> gpurun_out/r2c8_bench.err
# configs[1] shape: 100 Mb reference, paired-end 2 x 150 bp, 1 % error (2 M pairs here; generation time bounds the size)
timeout 1500 python tools/e2e_reads_bench.py --ref-len 100000000 --reads 2000000 --paired --threads 16 --repeat 2 \
   --binaries rabbitsalign_gasalgpu,rabbitsalign_gasalgpu_timed,rabbitsalign_b200_big,rabbitsalign_b200_gpuseed,rabbitsalign_b200_gpuseed_timed \
   > gpurun_out/r2c8_e2e_pe_2m.json 2> gpurun_out/r2c8_e2e_pe_2m.err
cat gpurun_out/r2c8_e2e_pe_2m.json | cut -c1-3000
tail -3 gpurun_out/r2c8_e2e_pe_2m.err
