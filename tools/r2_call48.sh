#!/bin/bash
# round 2, call 48: the bench's pipeline block alone, three runs per build, veneer timers on (call 47 saw 5 s for gpusam
# where call 45 saw 1.4 s)
cd /root/repo
mkdir -p gpurun_out
export RSA_EXT_STATS=1
timeout 900 python tools/e2e_reads_bench.py --ref-len 20000000 --reads 600000 --threads $(nproc) --repeat 3 \
  --binaries rabbitsalign_gasalgpu,rabbitsalign_b200_big,rabbitsalign_b200_gpuseed,rabbitsalign_b200_gpusam > gpurun_out/r2c48_e2e.json 2> gpurun_out/r2c48_e2e.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2c48_e2e.json'))
for k,v in d.items():
    if isinstance(v,dict): print(k, v.get('wall_s_runs'), v.get('mapping_s'), v.get('sam_md5'), v.get('error'), v.get('veneer_stats'))
PY
