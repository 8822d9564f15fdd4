mkdir -p gpurun_out
timeout 300 ./tools/microbench/read_bw > gpurun_out/read_bw.txt 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench8_2gpu.json 2> gpurun_out/bench8_2gpu.err; echo "rc=$?" >> gpurun_out/bench8_2gpu.err
