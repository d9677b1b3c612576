#!/bin/bash
# round 2, call 41: batched target staging + table-driven base codes in fast_dp_kernel: parity, then A/B against the previous
# build (rabbitsalign_b200/librsa_ext_head.so) on the same box, alternating
cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py -m gpu -x -q > gpurun_out/r2c41_pytest.txt 2>&1
tail -4 gpurun_out/r2c41_pytest.txt
for rep in 1 2; do
  for lib in head new; do
    if [ $lib = head ]; then export RSA_EXT_LIB=$PWD/rabbitsalign_b200/librsa_ext_head.so; else unset RSA_EXT_LIB; fi
    timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra-legs > gpurun_out/r2c41_bench_${lib}_$rep.json 2> gpurun_out/r2c41.err
    python -c "
import json
d=json.load(open('gpurun_out/r2c41_bench_${lib}_$rep.json')); print('$lib $rep value', round(d['value']), 'e2e', round(d['e2e']['value']), 'dp', round(d['roofline']['achieved']), 'win', round(d['detail']['e2e_windows_in_resident_reference']['gcups']) if 'e2e_windows_in_resident_reference' in d['detail'] else None, d['detail']['resident_equals_e2e_records'], d['detail']['records_sane'])"
  done
done
