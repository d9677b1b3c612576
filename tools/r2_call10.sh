#!/bin/bash
# round 2, GPU call 10 (2 GPUs): pipeline workers spread over two GPUs (SAM identity, reads/s), bench at N=2
set -x
cd /root/repo
nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_sam.py -m gpu -q -k "two-gpus or gpuseed" > gpurun_out/r2c10_pytest_sam_2gpu.txt 2>&1
tail -6 gpurun_out/r2c10_pytest_sam_2gpu.txt
for nd in 1 2; do
  RSA_EXT_DEVICES=$nd timeout 600 python tools/e2e_reads_bench.py --ref-len 50000000 --reads 3000000 --threads 16 --repeat 2 \
     --binaries rabbitsalign_b200_gpuseed > gpurun_out/r2c10_e2e_se_3m_dev$nd.json 2> gpurun_out/r2c10_e2e.err
  cat gpurun_out/r2c10_e2e_se_3m_dev$nd.json | cut -c1-900
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2c10_bench_n2.json 2> gpurun_out/r2c10_bench_n2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2c10_bench_n2.json').read().strip().splitlines()[-1])
print("N=2 value", round(d['value']), "e2e", round(d['e2e']['value']), d['n_gpus'])
PY
