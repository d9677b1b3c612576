#!/bin/bash
# round 2, call 13: A/B builds (lanes per group, register cap) + ncu launch list and full capture of the new DP kernel
set -x
cd /root/repo
mkdir -p gpurun_out
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra-legs"
show() { python - "$1" <<'PY'
import json,sys
d=json.load(open(sys.argv[1]))
print(sys.argv[1], "value", round(d['value']), "e2e", round(d['e2e']['value']), "dp", round(d['roofline']['achieved']), "tb_ms", round(d['roofline'].get('tb_ms',0),2))
PY
}
timeout 600 $B > gpurun_out/r2c13_bench_default.json 2> gpurun_out/r2c13.err; show gpurun_out/r2c13_bench_default.json
RSA_EXT_LIB=/root/repo/variants/librsa_ext_l8.so timeout 600 $B > gpurun_out/r2c13_bench_l8.json 2>> gpurun_out/r2c13.err; show gpurun_out/r2c13_bench_l8.json
RSA_EXT_LIB=/root/repo/variants/librsa_ext_w3.so timeout 600 $B > gpurun_out/r2c13_bench_w3.json 2>> gpurun_out/r2c13.err; show gpurun_out/r2c13_bench_w3.json
timeout 600 $B --read-len 250 --pairs 524288 > gpurun_out/r2c13_bench_250.json 2>> gpurun_out/r2c13.err; show gpurun_out/r2c13_bench_250.json
timeout 600 $B --read-len 100 --pairs 524288 > gpurun_out/r2c13_bench_100.json 2>> gpurun_out/r2c13.err; show gpurun_out/r2c13_bench_100.json
S="python bench.py --pairs 262144 --steps 2 --warmup 3 --no-cpu-baseline --no-extra-legs"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2c13_launches.csv $S > gpurun_out/r2c13_ncu1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:fast_dp_kernel|tb_groups_kernel' -c 4 -o gpurun_out/r2c13_prof $S > gpurun_out/r2c13_ncu2.log 2>&1
ls -la gpurun_out | tail -12
