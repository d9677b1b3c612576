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
  echo "$tag wall $(( (e - s) / 1000000 )) ms; $(grep -a 'Total time mapping' gpurun_out/pipe_$tag.err); $(grep -a 'Total time indexing' gpurun_out/pipe_$tag.err)"
}
for i in 1 2 3; do run b200_$i rabbitsalign_b200 RSA_EXT_TRACE=1; done
run b200_nowarm rabbitsalign_b200 RSA_EXT_TRACE=1 RSA_EXT_NO_WARMUP=1
for i in 1 2; do run cpu_$i rabbitsalign_cpussw X=1; done
