#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/b32_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/b32_pytest.log
timeout 1500 python bench.py > gpurun_out/b32_bench_n1.json 2> gpurun_out/b32_bench_n1.err; echo "rc=$?" >> gpurun_out/b32_bench_n1.err
timeout 600 python bench.py --impl reference > gpurun_out/b32_bench_ref.json 2> gpurun_out/b32_bench_ref.err; echo "rc=$?" >> gpurun_out/b32_bench_ref.err
tail -3 gpurun_out/b32_pytest.log; tail -c 300 gpurun_out/b32_bench_n1.err; cut -c1-400 gpurun_out/b32_bench_n1.json; cut -c1-600 gpurun_out/b32_bench_ref.json
