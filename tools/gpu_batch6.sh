#!/bin/bash
# 8 GPUs: the scaling run (headline only at each N), then the full N=1 line
mkdir -p gpurun_out
for N in 8 4; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 5 --warmup 3 --skip-extras > gpurun_out/b6_bench_n$N.json 2> gpurun_out/b6_bench_n$N.err; echo "rc=$?" >> gpurun_out/b6_bench_n$N.err
done
NCCL_DEBUG=INFO timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 2 --warmup 3 --skip-extras > gpurun_out/b6_bench_n8_dbg.json 2> gpurun_out/b6_bench_n8_dbg.err
grep -i "nvls\|algo\|channels" gpurun_out/b6_bench_n8_dbg.err | head -20 > gpurun_out/b6_nccl_info.txt
tail -c 400 gpurun_out/b6_bench_n8.err; tail -c 300 gpurun_out/b6_bench_n4.err
