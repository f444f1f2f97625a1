#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/r2_probe.py --scale 0.125 --what spmv,bb --panel-mb 48 > gpurun_out/b9_probe_s8_v8.log 2>&1
BSLS_SPMV_V8=0 timeout 300 python tools/r2_probe.py --scale 0.125 --what spmv,bb --panel-mb 48 > gpurun_out/b9_probe_s8_v4.log 2>&1
timeout 600 python tools/r2_probe.py --scale 1 --what spmv,bb --panel-mb 48 > gpurun_out/b9_probe_s1_v8.log 2>&1
timeout 1500 python bench.py > gpurun_out/b9_bench_n1.json 2> gpurun_out/b9_bench_n1.err; echo "rc=$?" >> gpurun_out/b9_bench_n1.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:solver_tiny --launch-skip 1 -c 1 -o gpurun_out/b9_tiny python tools/r2_probe.py --what c1 > gpurun_out/b9_ncu_tiny.log 2>&1
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/b9_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/b9_pytest.log
tail -3 gpurun_out/b9_pytest.log; tail -c 300 gpurun_out/b9_bench_n1.err
