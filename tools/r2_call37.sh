#!/bin/bash
# round 2, call 37: run-to-run spread of the gpuseed and gpuham pipeline builds on one data set (3 M x 150 bp SE, 50 Mb),
# alternating, with the veneer's timers (RSA_EXT_STATS=1); then BASELINE configs[4] at its full size through bench.py
cd /root/repo
mkdir -p gpurun_out
D=/tmp/r2c37; mkdir -p $D
python tools/make_reads.py $D --ref-len 50000000 --contigs 4 --reads 3000000 --seed 77 --read-len 150 --sub 0.01 --indel 0.003 --max-indel 3
export RSA_EXT_STATS=1
: > gpurun_out/r2c37_runs.txt
for rep in 1 2 3 4; do
  for b in rabbitsalign_b200_gpuseed rabbitsalign_b200_gpuham; do
    s=$(date +%s%N)
    integration/_build/$b -t $(nproc) -o $D/out.sam $D/ref.fa $D/reads_1.fq 2> $D/err.txt
    e=$(date +%s%N)
    echo "== $b run $rep wall_ms $(( (e - s) / 1000000 )) md5 $(grep -v '^@PG' $D/out.sam | md5sum | cut -c1-12)" >> gpurun_out/r2c37_runs.txt
    grep -h "Total time mapping\|base level\|rsa_ext veneer\|rsa_seed" $D/err.txt >> gpurun_out/r2c37_runs.txt
  done
done
cat gpurun_out/r2c37_runs.txt | grep -v "rsa_seed" | cut -c1-230
unset RSA_EXT_STATS
timeout 1500 python bench.py --pairs 10485760 --steps 15 --warmup 3 --no-cpu-baseline --no-extra-legs > gpurun_out/r2c35_bench_10m.json 2> gpurun_out/r2c35.err
tail -3 gpurun_out/r2c35.err
python -c "
import json
d=json.load(open('gpurun_out/r2c35_bench_10m.json')); print('value', round(d['value']), 'e2e', round(d['e2e']['value']), 'dp', round(d['roofline']['achieved']), 'ms/step', round(d['ms_per_step'],1), d['detail']['resident_equals_e2e_records'], d['detail']['records_sane'], d['clocks'])"
