#!/bin/bash
# round 2, call 42: the whole GPU suite on the final tree, the all-kernel-families probe, the deferral counters of a paired-end
# run, and BASELINE configs[1] at scale (100 Mb, 5 M pairs) through the reference's GPU build, the seeding build and the full
# device path (gpusam)
cd /root/repo
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/r2c42_pytest.txt 2>&1
tail -4 gpurun_out/r2c42_pytest.txt
timeout 600 python tools/all_kernels_probe.py > gpurun_out/r2c42_probe.txt 2>&1; tail -5 gpurun_out/r2c42_probe.txt
# paired-end, RabbitFX build: how many candidates took the device shortcut (deferral starts at 400 insert-size samples per chunk)
D=/tmp/r2c42; mkdir -p $D
python tools/make_reads.py $D --ref-len 2000000 --contigs 4 --reads 50000 --paired --seed 13
RSA_EXT_STATS=1 integration/_build/rabbitsalign_fx_b200_gpusam -t 4 -o $D/o.sam $D/ref.fa $D/reads_1.fq $D/reads_2.fq 2>&1 | grep "rsa_ext veneer" > gpurun_out/r2c42_pe_deferral.txt
cat gpurun_out/r2c42_pe_deferral.txt
timeout 2400 python tools/e2e_reads_bench.py --ref-len 100000000 --reads 5000000 --paired --threads $(nproc) \
  --binaries rabbitsalign_gasalgpu,rabbitsalign_b200_gpuseed,rabbitsalign_b200_gpusam > gpurun_out/r2c42_e2e_pe_5m.json 2> gpurun_out/r2c42_e2e.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2c42_e2e_pe_5m.json'))
for k,v in d.items():
    if isinstance(v,dict): print(k, v.get('wall_s'), v.get('mapping_s'), v.get('reads_per_s_wall'), v.get('sam_md5'), v.get('error'))
    else: print(k,v)
PY
