#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/r2_probe.py --scale 0.125 --what spmv --panel-mb 27,32,40,54,64,80,100,160 > gpurun_out/b23_panels_s8.log 2>&1
timeout 900 python tools/r2_probe.py --scale 0.25 --what spmv --panel-mb 40,48,54,64,80 > gpurun_out/b23_panels_s4.log 2>&1
grep A_x gpurun_out/b23_panels_s8.log | cut -c1-220; grep A_x gpurun_out/b23_panels_s4.log | cut -c1-220
