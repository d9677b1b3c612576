#!/bin/bash
# round 2, call 18: the whole GPU suite on the current tree + BASELINE configs[1] at scale (100 Mb reference, 5 M read
# pairs = 10 M reads of 150 bp, 1 % error) through the pipeline builds
cd /root/repo
mkdir -p gpurun_out
nproc; df -h /tmp | tail -1
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/r2c18_pytest.txt 2>&1
tail -8 gpurun_out/r2c18_pytest.txt
timeout 2400 python tools/e2e_reads_bench.py --ref-len 100000000 --reads 5000000 --paired --threads $(nproc) \
  --binaries rabbitsalign_gasalgpu,rabbitsalign_b200_big,rabbitsalign_b200_win,rabbitsalign_b200_gpuseed > gpurun_out/r2c18_e2e_pe_5m.json 2> gpurun_out/r2c18_e2e.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2c18_e2e_pe_5m.json'))
for k,v in d.items():
    if isinstance(v,dict): print(k, v.get('wall_s'), v.get('mapping_s'), v.get('reads_per_s_wall'), v.get('error'))
    else: print(k,v)
PY
