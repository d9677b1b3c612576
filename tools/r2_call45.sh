#!/bin/bash
# round 2, call 45 (final tree): the round's recorded measurements -- bench line with all legs, reference arm, read-length legs,
# microbenchmark, ncu launch list + full capture of the DP and traceback kernels
cd /root/repo
mkdir -p gpurun_out
T0=$(date +%s); timeout 1200 python bench.py > gpurun_out/r2c45_bench.json 2> gpurun_out/r2c45_bench.err; echo "default bench.py run: $(( $(date +%s) - T0 )) s"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2c45_bench.json'))
print("value", round(d['value']), "e2e", round(d['e2e']['value']), "dp", round(d['roofline']['achieved']), round(d['roofline']['frac'],3), "cpu", d['cpu_baseline'])
for k in ('leg_250bp_5pct_indel','hamming_shortcut','sam_format','pipeline'):
    print(k, json.dumps(d['detail'].get(k))[:600])
print('seeding', {k:v for k,v in (d['detail'].get('seeding') or {}).items() if k in ('reads_per_s_resident','reads_per_s_e2e','cpu_baseline','error')})
PY
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2c45_bench_ref.json 2>/dev/null; head -c 400 gpurun_out/r2c45_bench_ref.json; echo
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra-legs --pairs 524288"
for rl in 250 300; do timeout 600 $B --read-len $rl > gpurun_out/r2c45_bench_$rl.json 2>>gpurun_out/r2c45_bench.err; python -c "
import json,sys
d=json.load(open('gpurun_out/r2c45_bench_$rl.json')); print($rl, 'value', round(d['value']), 'e2e', round(d['e2e']['value']), 'dp', round(d['roofline']['achieved']))"; done
timeout 300 ./tools/dpx_microbench > gpurun_out/r2c45_dpx_microbench.jsonl 2>&1
S="python bench.py --pairs 262144 --steps 2 --warmup 3 --no-cpu-baseline --no-extra-legs"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2c45_launches.csv $S > gpurun_out/r2c45_ncu1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:fast_dp_kernel|tb_groups_kernel' -c 4 -o gpurun_out/r2c45_prof $S > gpurun_out/r2c45_ncu2.log 2>&1
ls -la gpurun_out | tail -6
