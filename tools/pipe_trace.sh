# diagnostic: run the integrated pipeline binaries a few times on one synthetic read set with RSA_EXT_TRACE=1
# usage: bash tools/pipe_trace.sh [ref_len] [reads] [threads]
REF=${1:-50000000}; READS=${2:-400000}; T=${3:-16}
D=/tmp/rd; mkdir -p $D gpurun_out
python tools/make_reads.py $D --ref-len $REF --contigs 4 --reads $READS --seed 77 > /dev/null
B=integration/_build
run() {  # tag, binary, env...
  tag=$1; exe=$2; shift 2
  s=$(date +%s%N)
  env "$@" $B/$exe -t $T -o $D/o.sam $D/ref.fa $D/reads_1.fq 2> gpurun_out/pipe_$tag.err
  e=$(date +%s%N)
  f=gpurun_out/pipe_$tag.err
  echo "$tag wall $(( (e - s) / 1000000 )) ms; $(grep -a 'Total time mapping' $f); $(grep -a 'Total time indexing' $f); ctx $(grep -a -m1 'context' $f | awk '{print $2, $(NF-1)}'); first chunk $(grep -a -m1 ' chunk ' $f | awk '{print $2}'); last $(grep -a 'retire\|chunk' $f | tail -1 | awk '{print $2}'); calls $(grep -a -c 'retire' $f)"
}
for i in 1 2 3; do run b200_$i rabbitsalign_b200 RSA_EXT_TRACE=1; done
for i in 1 2 3; do run sync_$i rabbitsalign_b200 RSA_EXT_TRACE=1 RSA_EXT_WARMUP=sync; done
for i in 1 2; do run aln_$i rabbitsalign_b200_alninfo RSA_EXT_TRACE=1; done
for i in 1 2; do run alnsync_$i rabbitsalign_b200_alninfo RSA_EXT_TRACE=1 RSA_EXT_WARMUP=sync; done
run cpu_1 rabbitsalign_cpussw X=1
