#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/r2_probe.py --what c1,c4 > gpurun_out/b10_probe_small.log 2>&1
timeout 300 python tools/r2_probe.py --scale 0.125 --what bb > gpurun_out/b10_probe_s8.log 2>&1
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/b10_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/b10_pytest.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:solver_tiny --launch-skip 1 -c 1 -o gpurun_out/b10_tiny python tools/r2_probe.py --what c1 > gpurun_out/b10_ncu_tiny.log 2>&1
tail -3 gpurun_out/b10_pytest.log; cat gpurun_out/b10_probe_small.log | cut -c1-250
