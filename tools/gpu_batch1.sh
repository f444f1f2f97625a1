#!/bin/bash
# round-2 GPU batch 1: parity suite + SpMV / solver-loop probes
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/b1_smi.txt 2>&1
nproc >> gpurun_out/b1_smi.txt; free -g >> gpurun_out/b1_smi.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/b1_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/b1_pytest.log
timeout 600 python tools/r2_probe.py --scale 0.125 --what spmv,bb,c1,c4 --panel-mb 32,48,64 > gpurun_out/b1_probe_s8.log 2>&1
BSLS_ELL_TEX=1 timeout 300 python tools/r2_probe.py --scale 0.125 --what spmv --panel-mb 48 > gpurun_out/b1_probe_s8_tex.log 2>&1
BSLS_BATCH_LEGACY=1 timeout 600 python tools/r2_probe.py --scale 0.125 --what bb,c1,c4 > gpurun_out/b1_probe_s8_legacy.log 2>&1
timeout 900 python tools/r2_probe.py --scale 1 --what spmv,bb --panel-mb 48 > gpurun_out/b1_probe_s1.log 2>&1
BSLS_BATCH_LEGACY=1 timeout 600 python tools/r2_probe.py --scale 1 --what bb > gpurun_out/b1_probe_s1_legacy.log 2>&1
tail -3 gpurun_out/b1_pytest.log
