# bench.py at N = 2, 4, 8 on one box (launched the way the driver does); output under gpurun_out/
for N in 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + N)) \
    bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
  tail -c 600 gpurun_out/bench_n$N.json | head -c 300; echo
done
