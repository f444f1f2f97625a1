#!/bin/bash
N=$1
mkdir -p gpurun_out
BSLS_P2P_PROF=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus $N --steps 5 --warmup 3 --skip-extras > gpurun_out/b33_bench_n$N.json 2> gpurun_out/b33_bench_n$N.err; echo "rc=$?" >> gpurun_out/b33_bench_n$N.err
grep "p2p prof" gpurun_out/b33_bench_n$N.err; tail -n 2 gpurun_out/b33_bench_n$N.err; cut -c1-120 gpurun_out/b33_bench_n$N.json
