#!/bin/bash
# round 2, GPU call 3: e2e after dropping the upfront validation pass; chunk ramp variants
set -x
cd /root/repo
timeout 600 python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -5
for ramp in default 32768,65536 65536 131072 16384,131072 8192,32768,131072; do
  if [ "$ramp" = default ]; then unset RSA_EXT_RAMP; else export RSA_EXT_RAMP=$ramp; fi
  timeout 300 python bench.py --steps 10 --warmup 3 --no-extra-legs --no-cpu-baseline > gpurun_out/r2c3_bench_$ramp.json 2> gpurun_out/r2c3_bench_$ramp.err
  python - "$ramp" <<'PY'
import json,sys
d=json.load(open('gpurun_out/r2c3_bench_%s.json'%sys.argv[1]))
print("RAMP", sys.argv[1], "value", round(d['value']), "e2e", round(d['e2e']['value']), "ratio", round(d['e2e']['value']/d['value'],3), "plan_ms", round(d['e2e']['host_plan_ms_per_step'],2), "ms", round(d['e2e']['ms_per_step'],2), "roofline", d['roofline'].get('frac'), d['roofline'].get('peak'))
PY
done
unset RSA_EXT_RAMP
RSA_EXT_TRACE=1 timeout 120 python tools/e2e_timeline.py 2>&1 | grep -v create | tail -45
