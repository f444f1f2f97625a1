#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/r2_probe.py --what c1,c4 > gpurun_out/b13_probe_small.log 2>&1
timeout 1000 python -m pytest tests -m gpu -q -x > gpurun_out/b13_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/b13_pytest.log
tail -3 gpurun_out/b13_pytest.log; cut -c1-250 gpurun_out/b13_probe_small.log
