#!/bin/bash
# round 2, call 44: pinned bounce buffers for pageable callers of the Hamming and SAM entry points: parity of both, then
# BASELINE configs[1] at scale (tmpfs) for the seeding build and the full device path
cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_hamming.py tests/test_gpu_sam_format.py tests/test_gpu_sam.py -m gpu -q -k "hamming or format or gpusam or gpuham" > gpurun_out/r2c44_pytest.txt 2>&1
tail -4 gpurun_out/r2c44_pytest.txt
export RSA_EXT_STATS=1
timeout 2400 python tools/e2e_reads_bench.py --ref-len 100000000 --reads 5000000 --paired --threads $(nproc) --repeat 2 \
  --binaries rabbitsalign_b200_gpuseed,rabbitsalign_b200_gpuham,rabbitsalign_b200_gpusam > gpurun_out/r2c44_e2e_pe_5m.json 2> gpurun_out/r2c44_e2e.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2c44_e2e_pe_5m.json'))
for k,v in d.items():
    if isinstance(v,dict): print(k, v.get('wall_s_runs'), v.get('mapping_s'), v.get('reads_per_s_wall'), v.get('sam_md5'), v.get('error'), v.get('veneer_stats'))
    else: print(k,v)
PY
