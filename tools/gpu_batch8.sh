#!/bin/bash
mkdir -p gpurun_out
for N in 8 4; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2961$N bench.py --gpus $N --steps 5 --warmup 3 --skip-extras > gpurun_out/b8_bench_n${N}_p2p.json 2> gpurun_out/b8_bench_n${N}_p2p.err; echo "rc=$?" >> gpurun_out/b8_bench_n${N}_p2p.err
done
BSLS_P2P=0 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29631 bench.py --gpus 8 --steps 5 --warmup 3 --skip-extras > gpurun_out/b8_bench_n8_nccl.json 2> gpurun_out/b8_bench_n8_nccl.err; echo "rc=$?" >> gpurun_out/b8_bench_n8_nccl.err
tail -c 300 gpurun_out/b8_bench_n8_p2p.err
