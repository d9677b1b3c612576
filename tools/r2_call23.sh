#!/bin/bash
# round 2, call 23 (8 GPUs): N = 8 with and without the per-rank core slices; GPU/CPU topology for the record
cd /root/repo
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2c23_topo.txt 2>&1; lscpu | grep -i "numa\|socket\|model name" > gpurun_out/r2c23_lscpu.txt
head -12 gpurun_out/r2c23_topo.txt | cut -c1-160; cat gpurun_out/r2c23_lscpu.txt
for tag in aff noaff; do
  if [ $tag = noaff ]; then export RSA_BENCH_NO_AFFINITY=1; else unset RSA_BENCH_NO_AFFINITY; fi
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 \
    bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2c23_bench_n8_$tag.json 2> gpurun_out/r2c23_n8_$tag.err
  python -c "
import json
d=json.loads(open('gpurun_out/r2c23_bench_n8_$tag.json').read().strip().splitlines()[-1])
print('$tag', 'value', round(d['value']), 'e2e', round(d['e2e']['value']), d['detail'].get('rank0_cpu_affinity'), d['detail']['e2e_windows_in_resident_reference'])"
done
