#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/pava_1e8.py > gpurun_out/b18_pava_1e8_fast.log 2>&1
BSLS_PAVA_EXACT=1 timeout 600 python tools/pava_1e8.py ref zspace > gpurun_out/b18_pava_1e8_exact.log 2>&1
timeout 300 python tools/c3_run.py > gpurun_out/b18_c3.log 2>&1
timeout 1200 python -m pytest tests/test_pava_gpu.py tests/test_fuzz_gpu.py tests/test_solvers_gpu.py tests/test_pipeline_gpu.py -m gpu -q -x > gpurun_out/b18_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/b18_pytest.log
cut -c1-260 gpurun_out/b18_pava_1e8_fast.log; echo; cut -c1-260 gpurun_out/b18_pava_1e8_exact.log; tail -1 gpurun_out/b18_c3.log | cut -c1-900; tail -8 gpurun_out/b18_pytest.log
