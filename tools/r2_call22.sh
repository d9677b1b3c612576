#!/bin/bash
# round 2, call 22 (8 GPUs): bench.py at N = 1, 2, 4, 8 on one box, launched the way the driver does
cd /root/repo
mkdir -p gpurun_out
nvidia-smi -L | wc -l; nproc
timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-extra-legs --no-cpu-baseline > gpurun_out/r2c22_bench_n1.json 2> gpurun_out/r2c22_n1.err
for N in 2 4 8; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + N)) \
    bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2c22_bench_n$N.json 2> gpurun_out/r2c22_n$N.err
done
python - <<'PY'
import json
base=None
for N in (1,2,4,8):
    try:
        d=json.loads(open(f'gpurun_out/r2c22_bench_n{N}.json').read().strip().splitlines()[-1])
    except Exception as e:
        print(N, 'FAILED', e); continue
    if N==1: base=(d['value'], d['e2e']['value'])
    print(N, 'value', round(d['value']), 'e2e', round(d['e2e']['value']), 'eff', round(d['value']/base[0]/N,3), round(d['e2e']['value']/base[1]/N,3), 'host_plan_ms', round(d['e2e'].get('host_plan_ms_per_step',0),2), d['clocks'].get('reasons'))
PY
