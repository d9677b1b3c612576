#!/bin/bash
# round 2, call 16: Hamming shortcut kernel parity + full GPU suite (without the SAM pipeline builds)
cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_hamming.py -m gpu -x -q > gpurun_out/r2c16_pytest_hamming.txt 2>&1
tail -15 gpurun_out/r2c16_pytest_hamming.txt
