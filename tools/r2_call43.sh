#!/bin/bash
# round 2, call 43: BASELINE configs[1] at scale again with inputs and SAM on tmpfs (call 42 wrote 3.5 GB of SAM per binary to
# the box's scratch disk and timed the disk: gpuseed 11.3 s against 6.5 s on the box of call 18)
cd /root/repo
mkdir -p gpurun_out
df -h /tmp /dev/shm | tail -2; free -g | head -2
export RSA_EXT_STATS=1
timeout 2400 python tools/e2e_reads_bench.py --ref-len 100000000 --reads 5000000 --paired --threads $(nproc) \
  --binaries rabbitsalign_gasalgpu,rabbitsalign_b200_gpuseed,rabbitsalign_b200_gpusam > gpurun_out/r2c43_e2e_pe_5m.json 2> gpurun_out/r2c43_e2e.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2c43_e2e_pe_5m.json'))
for k,v in d.items():
    if isinstance(v,dict): print(k, v.get('wall_s'), v.get('mapping_s'), v.get('reads_per_s_wall'), v.get('sam_md5'), v.get('error'), v.get('veneer_stats'))
    else: print(k,v)
PY
