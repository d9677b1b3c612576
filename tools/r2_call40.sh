#!/bin/bash
# round 2, call 40: the bare cell recipe at the product kernel's shape (38 columns per lane) and occupancy (3 warps per
# scheduler), and at higher occupancies -- does the recipe or the rest of the kernel limit fast_dp_kernel<4,38>?
cd /root/repo
mkdir -p gpurun_out
tools/dpx_microbench > gpurun_out/r2c40_micro.jsonl 2>&1
grep "bare cell recipe\|cell recipe" gpurun_out/r2c40_micro.jsonl
