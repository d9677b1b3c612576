#!/bin/bash
# round 2, call 38: SAM text from the device inside the pipeline (gpusam builds): byte-identical SAM on the golden configs,
# then run-to-run walls of gpuham vs gpusam on one data set (3 M x 150 bp SE, 50 Mb), alternating
cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sam.py -m gpu -q -k "gpusam or two-gpus" > gpurun_out/r2c38_pytest.txt 2>&1
tail -15 gpurun_out/r2c38_pytest.txt
D=/tmp/r2c38; mkdir -p $D
python tools/make_reads.py $D --ref-len 50000000 --contigs 4 --reads 3000000 --seed 77 --read-len 150 --sub 0.01 --indel 0.003 --max-indel 3
export RSA_EXT_STATS=1
: > gpurun_out/r2c38_runs.txt
for rep in 1 2 3; do
  for b in rabbitsalign_gasalgpu rabbitsalign_b200_gpuseed rabbitsalign_b200_gpuham rabbitsalign_b200_gpusam; do
    if [ $b = rabbitsalign_gasalgpu ] && [ $rep != 1 ]; then continue; fi
    s=$(date +%s%N)
    integration/_build/$b -t $(nproc) -o $D/out.sam $D/ref.fa $D/reads_1.fq 2> $D/err.txt
    e=$(date +%s%N)
    echo "== $b run $rep wall_ms $(( (e - s) / 1000000 )) md5 $(grep -v '^@PG' $D/out.sam | md5sum | cut -c1-12)" >> gpurun_out/r2c38_runs.txt
    grep -h "Total time mapping\|base level\|rsa_ext veneer" $D/err.txt >> gpurun_out/r2c38_runs.txt
  done
done
grep "==\|Total time mapping" gpurun_out/r2c38_runs.txt
