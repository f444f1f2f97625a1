#!/bin/bash
# usage: gpu_batch31.sh N : the driver's command line (no --skip-extras)
N=$1
mkdir -p gpurun_out
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/b31_bench_full_n$N.json 2> gpurun_out/b31_bench_full_n$N.err; echo "rc=$?" >> gpurun_out/b31_bench_full_n$N.err
tail -c 300 gpurun_out/b31_bench_full_n$N.err; cut -c1-200 gpurun_out/b31_bench_full_n$N.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/b31_ref_n$N.json 2> gpurun_out/b31_ref_n$N.err; echo "rc=$?" >> gpurun_out/b31_ref_n$N.err
tail -c 200 gpurun_out/b31_ref_n$N.err; cut -c1-200 gpurun_out/b31_ref_n$N.json
