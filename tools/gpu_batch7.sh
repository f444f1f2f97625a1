#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_dist_gpu.py -q > gpurun_out/b7_pytest_dist.log 2>&1; echo "rc=$?" >> gpurun_out/b7_pytest_dist.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 3 --warmup 3 --skip-extras > gpurun_out/b7_bench_n2_p2p.json 2> gpurun_out/b7_bench_n2_p2p.err; echo "rc=$?" >> gpurun_out/b7_bench_n2_p2p.err
BSLS_P2P=0 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 2 --steps 3 --warmup 3 --skip-extras > gpurun_out/b7_bench_n2_nccl.json 2> gpurun_out/b7_bench_n2_nccl.err; echo "rc=$?" >> gpurun_out/b7_bench_n2_nccl.err
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/b7_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/b7_pytest.log
tail -15 gpurun_out/b7_pytest_dist.log; tail -c 600 gpurun_out/b7_bench_n2_p2p.err; tail -3 gpurun_out/b7_pytest.log
