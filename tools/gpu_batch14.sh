#!/bin/bash
mkdir -p gpurun_out
BSLS_TINY_PROF=1 timeout 300 python tools/r2_probe.py --what c1 > gpurun_out/b14_probe.log 2>&1
cut -c1-400 gpurun_out/b14_probe.log
