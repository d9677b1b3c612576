#!/bin/bash
# round 2, call 36: the Hamming shortcut wired into the pipeline (gpuham builds): byte-identical SAM on the golden configs,
# then reads/s against the gpuseed build and the reference's GPU build (150 bp SE 3 M; PE 2 M pairs, 100 Mb)
cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sam.py -m gpu -q -k "gpuham or gpuseed" > gpurun_out/r2c36_pytest.txt 2>&1
tail -8 gpurun_out/r2c36_pytest.txt
export RSA_EXT_STATS=1
timeout 900 python tools/e2e_reads_bench.py --ref-len 50000000 --reads 3000000 --threads $(nproc) --repeat 2 \
  --binaries rabbitsalign_gasalgpu,rabbitsalign_b200_gpuseed,rabbitsalign_b200_gpuham > gpurun_out/r2c36_e2e_se_3m.json 2> gpurun_out/r2c36_e2e.err
timeout 900 python tools/e2e_reads_bench.py --ref-len 100000000 --reads 2000000 --paired --threads $(nproc) --repeat 2 \
  --binaries rabbitsalign_gasalgpu,rabbitsalign_b200_gpuseed,rabbitsalign_b200_gpuham > gpurun_out/r2c36_e2e_pe_2m.json 2>> gpurun_out/r2c36_e2e.err
python - <<'PY'
import json
for f in ('gpurun_out/r2c36_e2e_se_3m.json','gpurun_out/r2c36_e2e_pe_2m.json'):
    try:
        d=json.load(open(f)); print(f)
        for k,v in d.items():
            if isinstance(v,dict): print(' ', k, v.get('wall_s'), v.get('mapping_s'), v.get('reads_per_s_wall'), v.get('sam_md5'), v.get('error'), [t for t in v.get('stderr_times',[]) if 'base level' in t], v.get('veneer_stats'))
    except Exception as e: print(f, e)
PY
