#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/b2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/b2_pytest.log
timeout 300 python tools/r2_probe.py --what c1,c4 > gpurun_out/b2_probe_small.log 2>&1
timeout 1200 python bench.py > gpurun_out/b2_bench_n1.json 2> gpurun_out/b2_bench_n1.err; echo "bench rc=$?" >> gpurun_out/b2_bench_n1.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:spmv -c 8 -o gpurun_out/b2_spmv python tools/r2_probe.py --scale 0.125 --what spmv --panel-mb 48 > gpurun_out/b2_ncu.log 2>&1
tail -3 gpurun_out/b2_pytest.log; tail -c 600 gpurun_out/b2_bench_n1.err
