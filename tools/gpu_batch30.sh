#!/bin/bash
mkdir -p gpurun_out
BSLS_TINY_PROF=1 timeout 120 python tools/r2_probe.py --what c1 > gpurun_out/b30_probe.log 2>&1; echo "rc=$?" >> gpurun_out/b30_probe.log
cut -c1-330 gpurun_out/b30_probe.log
timeout 900 python -m pytest tests/test_solvers_gpu.py -m gpu -q -x -k "batch or config1" > gpurun_out/b30_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/b30_pytest.log
tail -15 gpurun_out/b30_pytest.log
