#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/b4_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/b4_pytest.log
timeout 300 python tools/r2_probe.py --what c1,c4 > gpurun_out/b4_probe_small.log 2>&1
timeout 300 python tools/r2_probe.py --scale 0.125 --what spmv,bb --panel-mb 48 > gpurun_out/b4_probe_s8.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:solver_tiny -c 1 -o gpurun_out/b4_tiny python tools/r2_probe.py --what c1 > gpurun_out/b4_ncu_tiny.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:spmv_vector --launch-skip 3 -c 3 -o gpurun_out/b4_vec python tools/r2_probe.py --scale 0.125 --what spmv --panel-mb 48 > gpurun_out/b4_ncu_vec.log 2>&1
tail -5 gpurun_out/b4_pytest.log
