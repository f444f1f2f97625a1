#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/c3_run.py > gpurun_out/b24_c3.log 2>&1
timeout 600 python -m pytest tests/test_projection_gpu.py tests/test_fuzz_gpu.py -m gpu -q -x > gpurun_out/b24_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/b24_pytest.log
tail -1 gpurun_out/b24_c3.log | cut -c1-500; tail -3 gpurun_out/b24_pytest.log
