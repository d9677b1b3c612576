#!/bin/bash
# round 2, call 32: where does a whole-chunk call spend its time on 250-bp indel-rich reads? (RSA_EXT_STATS=1)
cd /root/repo
mkdir -p gpurun_out
D=/tmp/rd250; mkdir -p $D
python tools/make_reads.py $D --ref-len 50000000 --contigs 4 --reads 1000000 --seed 77 --read-len 250 --sub 0.02 --indel 0.02 --max-indel 4 > /dev/null
B=integration/_build
for rep in 1 2; do
for exe in rabbitsalign_gasalgpu rabbitsalign_b200_big rabbitsalign_b200; do
  s=$(date +%s%N)
  RSA_EXT_STATS=1 $B/$exe -t 16 -o $D/o.sam $D/ref.fa $D/reads_1.fq 2> gpurun_out/r2c32_$exe.err
  e=$(date +%s%N)
  echo "$exe wall $(( (e - s) / 1000000 )) ms; $(grep -a 'Total time mapping' gpurun_out/r2c32_$exe.err); $(grep -a 'rsa_ext veneer' gpurun_out/r2c32_$exe.err)"
done
done
