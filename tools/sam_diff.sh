# diagnostic: run every pipeline build on one synthetic read set and report the first differing SAM records
# usage: bash tools/sam_diff.sh <reads> <read_len> <sub> <indel> <max_indel> [threads] ["build list", first = baseline]
READS=${1:-100000}; RL=${2:-250}; SUB=${3:-0.02}; INDEL=${4:-0.02}; MI=${5:-4}; T=${6:-16}
BUILDS=${7:-"gasalref gasalgpu b200 b200_alninfo"}
PAIRED=${PAIRED:-0}  # PAIRED=1: paired-end reads
NRATE=${NRATE:-0}    # fraction of read bases replaced by N
PREFIX=${PREFIX:-rabbitsalign_}  # PREFIX=rabbitsalign_fx_ : the shipped RabbitFX configuration
D=/tmp/sd; rm -rf $D; mkdir -p $D gpurun_out
python tools/make_reads.py $D --ref-len 20000000 --contigs 4 --reads $READS --seed 77 --read-len $RL --sub $SUB --indel $INDEL --max-indel $MI --n-rate $NRATE $([ "$PAIRED" = 1 ] && echo --paired) > /dev/null
FQ="$D/reads_1.fq"; [ "$PAIRED" = 1 ] && FQ="$D/reads_1.fq $D/reads_2.fq"
B=integration/_build
for exe in $BUILDS; do
  $B/$PREFIX$exe -t $T -o $D/$exe.sam $D/ref.fa $FQ 2> $D/$exe.err || echo "$exe failed"
  grep -v '^@PG' $D/$exe.sam > $D/$exe.nopg
  echo "$exe $(md5sum < $D/$exe.nopg | cut -c1-8) $(wc -l < $D/$exe.nopg) lines"
done
set -- $BUILDS
base=$1; shift
for other in "$@"; do
  n=$(diff $D/$base.nopg $D/$other.nopg | grep -c '^<')
  echo "== $base vs $other: $n differing records"
  diff $D/$base.nopg $D/$other.nopg | head -8 | cut -c1-400
done
