python bench.py --steps 10 --warmup 3 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_final.json 2>/dev/null
python bench.py --read-len 300 --pairs 524288 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_300_final.json 2>/dev/null
python bench.py --read-len 250 --pairs 524288 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_250_final.json 2>/dev/null
tools/dpx_microbench > gpurun_out/micro_final.jsonl 2>&1
python bench.py --pairs 262144 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/b_small.log 2>&1 && ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_final.csv python bench.py --pairs 262144 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu1.log 2>&1
python bench.py --pairs 262144 --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k 'regex:fast_dp_kernel|tb_groups_kernel' -c 4 -o gpurun_out/prof_final python bench.py --pairs 262144 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu2.log 2>&1
ls -la gpurun_out/ | tail -12
