#!/bin/bash
# round 2, call 29: the whole GPU suite on the final tree, the bench line, pipeline runs (150 bp SE 3 M; 250 bp indel-rich 1 M)
cd /root/repo
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/r2c29_pytest.txt 2>&1
tail -5 gpurun_out/r2c29_pytest.txt
timeout 1200 python bench.py --steps 20 --warmup 3 > gpurun_out/r2c29_bench.json 2> gpurun_out/r2c29_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2c29_bench.json'))
print("value", round(d['value']), "e2e", round(d['e2e']['value']), "dp", round(d['roofline']['achieved']), round(d['roofline']['frac'],3))
print('indel leg', d['detail']['leg_250bp_5pct_indel'])
print('pipeline', json.dumps(d['detail']['pipeline'])[:900])
PY
timeout 900 python tools/e2e_reads_bench.py --ref-len 50000000 --reads 3000000 --threads $(nproc) --repeat 2 \
  --binaries rabbitsalign_gasalgpu,rabbitsalign_b200_big,rabbitsalign_b200_win,rabbitsalign_b200_gpuseed > gpurun_out/r2c29_e2e_se_3m.json 2> gpurun_out/r2c29_e2e.err
timeout 900 python tools/e2e_reads_bench.py --ref-len 50000000 --reads 1000000 --threads $(nproc) --repeat 2 --read-len 250 --sub 0.02 --indel 0.02 --max-indel 4 \
  --binaries rabbitsalign_gasalgpu,rabbitsalign_b200_big,rabbitsalign_b200_gpuseed > gpurun_out/r2c29_e2e_se_250bp.json 2>> gpurun_out/r2c29_e2e.err
python - <<'PY'
import json
for f in ('gpurun_out/r2c29_e2e_se_3m.json','gpurun_out/r2c29_e2e_se_250bp.json'):
    d=json.load(open(f)); print(f)
    for k,v in d.items():
        if isinstance(v,dict): print(' ', k, v.get('wall_s'), v.get('mapping_s'), v.get('reads_per_s_wall'), v.get('sam_md5'), v.get('error'))
PY
