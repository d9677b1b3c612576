# diagnostic: paired-end SAM of the reference pipeline is not reproducible run to run with many threads (its insert-size
# estimate depends on chunk timing) -- shown by running the SAME build twice; with -t 1 all builds agree byte for byte.
D=/tmp/sd2; rm -rf $D; mkdir -p $D
python tools/make_reads.py $D --ref-len 20000000 --contigs 4 --reads 1000000 --seed 77 --paired > /dev/null
B=integration/_build
for tag in gasalgpu.1 gasalgpu.2 b200.1 b200.2; do
  exe=${tag%.*}
  $B/rabbitsalign_$exe -t 16 -o $D/$tag.sam $D/ref.fa $D/reads_1.fq $D/reads_2.fq 2> /dev/null
  grep -v '^@PG' $D/$tag.sam > $D/$tag.nopg; rm $D/$tag.sam
  echo "$tag $(md5sum < $D/$tag.nopg | cut -c1-8)"
done
for p in "gasalgpu.1 gasalgpu.2" "b200.1 b200.2" "gasalgpu.1 b200.1"; do set -- $p; echo "== $1 vs $2: $(diff $D/$1.nopg $D/$2.nopg | grep -c '^<') differing lines"; done
# single-thread runs are deterministic by construction
for exe in gasalgpu b200; do $B/rabbitsalign_$exe -t 1 -o $D/$exe.t1.sam $D/ref.fa <(head -400000 $D/reads_1.fq) <(head -400000 $D/reads_2.fq) 2> /dev/null; grep -v '^@PG' $D/$exe.t1.sam | md5sum | cut -c1-8; done
