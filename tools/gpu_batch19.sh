#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/pava_1e8.py ref zspace normal > gpurun_out/b19_pava_1e8_fast.log 2>&1
BSLS_PAVA_EXACT=1 timeout 600 python tools/pava_1e8.py normal > gpurun_out/b19_pava_1e8_exact.log 2>&1
timeout 300 python tools/c3_run.py > gpurun_out/b19_c3.log 2>&1
cut -c1-260 gpurun_out/b19_pava_1e8_fast.log; echo; cut -c1-260 gpurun_out/b19_pava_1e8_exact.log; tail -1 gpurun_out/b19_c3.log | cut -c1-900
