#!/bin/bash
# round 2, call 25: parity after the IMAD fact merges, then the recorded measurements again (as tools/r2_call20.sh)
cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py tests/test_gpu_alninfo.py tests/test_gpu_reference_gpu.py tests/test_gpu_hamming.py tests/test_gpu_sam_format.py -m gpu -x -q > gpurun_out/r2c25_pytest.txt 2>&1
tail -4 gpurun_out/r2c25_pytest.txt
sed -i 's/r2c20/r2c25/g' tools/r2_call20.sh
bash tools/r2_call20.sh
