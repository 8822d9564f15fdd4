mkdir -p gpurun_out
run() { tag=$1; shift
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29536 bench.py --gpus 8 --steps 10 --warmup 3 --legs strong --no-cpu-baseline > gpurun_out/s8b_$tag.json 2> gpurun_out/s8b_$tag.err
}
run lanes2 BENCH_RESTART_LANES=2
run lanes1 BENCH_RESTART_LANES=1
