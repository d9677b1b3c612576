# diagnostic: run every pipeline build on one synthetic read set and report the first differing SAM records
# usage: bash tools/sam_diff.sh <reads> <read_len> <sub> <indel> <max_indel> [threads]
READS=${1:-100000}; RL=${2:-250}; SUB=${3:-0.02}; INDEL=${4:-0.02}; MI=${5:-4}; T=${6:-16}
D=/tmp/sd; rm -rf $D; mkdir -p $D gpurun_out
python tools/make_reads.py $D --ref-len 20000000 --contigs 4 --reads $READS --seed 77 --read-len $RL --sub $SUB --indel $INDEL --max-indel $MI > /dev/null
B=integration/_build
for exe in gasalref gasalgpu b200 b200_alninfo; do
  $B/rabbitsalign_$exe -t $T -o $D/$exe.sam $D/ref.fa $D/reads_1.fq 2> $D/$exe.err || echo "$exe failed"
  grep -v '^@PG' $D/$exe.sam > $D/$exe.nopg
  echo "$exe $(md5sum < $D/$exe.nopg | cut -c1-8) $(wc -l < $D/$exe.nopg) lines"
done
for pair in "gasalref gasalgpu" "gasalref b200" "b200 b200_alninfo"; do
  set -- $pair
  n=$(diff $D/$1.nopg $D/$2.nopg | grep -c '^<')
  echo "== $1 vs $2: $n differing records"
  diff $D/$1.nopg $D/$2.nopg | head -8 | cut -c1-400
done
cp $D/*.nopg gpurun_out/ 2>/dev/null; gzip -f gpurun_out/*.nopg
