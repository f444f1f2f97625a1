#!/bin/bash
# 2 GPUs: dist test, sharded bench (headline only), small probes on GPU 0
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_dist_gpu.py -q > gpurun_out/b5_pytest_dist.log 2>&1; echo "rc=$?" >> gpurun_out/b5_pytest_dist.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --skip-extras > gpurun_out/b5_bench_n2.json 2> gpurun_out/b5_bench_n2.err; echo "rc=$?" >> gpurun_out/b5_bench_n2.err
timeout 300 python tools/r2_probe.py --what c1 > gpurun_out/b5_probe_c1.log 2>&1
timeout 600 python bench.py --steps 3 --warmup 3 --skip-extras > gpurun_out/b5_bench_n1.json 2> gpurun_out/b5_bench_n1.err; echo "rc=$?" >> gpurun_out/b5_bench_n1.err
tail -3 gpurun_out/b5_pytest_dist.log; tail -c 1500 gpurun_out/b5_bench_n2.err; cat gpurun_out/b5_probe_c1.log | cut -c1-300
