#!/bin/bash
# round 2, call 30: 250-bp indel-rich pipeline with RSA_EXT_TRACE=1 (where does the time of a whole-chunk call go?)
cd /root/repo
mkdir -p gpurun_out
D=/tmp/rd250; mkdir -p $D
python tools/make_reads.py $D --ref-len 50000000 --contigs 4 --reads 1000000 --seed 77 --read-len 250 --sub 0.02 --indel 0.02 --max-indel 4 > /dev/null
B=integration/_build
for exe in rabbitsalign_gasalgpu rabbitsalign_b200_big; do
  s=$(date +%s%N)
  RSA_EXT_TRACE=1 $B/$exe -t 16 -o $D/o.sam $D/ref.fa $D/reads_1.fq 2> gpurun_out/r2c30_$exe.err
  e=$(date +%s%N)
  echo "$exe wall $(( (e - s) / 1000000 )) ms; $(grep -a 'Total time mapping' gpurun_out/r2c30_$exe.err)"
done
python - <<'PY'
import re,collections
f='gpurun_out/r2c30_rabbitsalign_b200_big.err'
laps=collections.defaultdict(list); ns=[]; t_first=None; t_last=None
for ln in open(f, errors='replace'):
    m=re.match(r'\[rsa_ext\s+([\d.]+) chunk (\S+) n=(\d+)\]\s+(.*?)\s+([\d.]+) ms', ln)
    if m:
        laps[m.group(4)].append(float(m.group(5))); 
        if m.group(4).startswith('plan'): ns.append(int(m.group(3)))
        t=float(m.group(1)); t_first=t if t_first is None else t_first; t_last=t
    m=re.match(r'\[rsa_ext\s+([\d.]+) create', ln)
for k,v in laps.items():
    v=sorted(v); print(f'{k:26s} n={len(v)} sum={sum(v):9.1f} ms median={v[len(v)//2]:.3f} p95={v[int(len(v)*0.95)]:.3f} max={v[-1]:.1f}')
print('calls', len(ns), 'pairs/call median', sorted(ns)[len(ns)//2] if ns else None, 'total pairs', sum(ns), 'first..last ms', t_first, t_last)
PY
grep -a "create" gpurun_out/r2c30_rabbitsalign_b200_big.err | head -5
