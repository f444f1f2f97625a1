#!/bin/bash
# usage: gpu_batch22.sh N   (run under gpurun --gpus N)
N=$1
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 --skip-extras > gpurun_out/b22_bench_n$N.json 2> gpurun_out/b22_bench_n$N.err; echo "rc=$?" >> gpurun_out/b22_bench_n$N.err
tail -c 400 gpurun_out/b22_bench_n$N.err; cut -c1-300 gpurun_out/b22_bench_n$N.json
