#!/bin/bash
# round 2, call 49: SAM formatter with one device arena (2 device + 2 pinned allocations on a first call instead of 7 + 3):
# parity, SAM identity of the gpusam builds, the bench's pipeline block (three runs per build)
cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_sam_format.py tests/test_gpu_sam.py -m gpu -q -k "format or text or gpusam" > gpurun_out/r2c49_pytest.txt 2>&1; tail -3 gpurun_out/r2c49_pytest.txt
export RSA_EXT_STATS=1
timeout 900 python tools/e2e_reads_bench.py --ref-len 20000000 --reads 600000 --threads $(nproc) --repeat 3 \
  --binaries rabbitsalign_gasalgpu,rabbitsalign_b200_gpuseed,rabbitsalign_b200_gpuham,rabbitsalign_b200_gpusam > gpurun_out/r2c49_e2e.json 2> gpurun_out/r2c49_e2e.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2c49_e2e.json'))
for k,v in d.items():
    if isinstance(v,dict): print(k, v.get('wall_s_runs'), v.get('mapping_s'), v.get('sam_md5'), v.get('error'), [x[16:120] for x in (v.get('veneer_stats') or [])])
PY
