#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"solver_cluster" -c 1 -o /tmp/b29_c1 python tools/c1_run.py > gpurun_out/b29_ncu.log 2>&1
ncu -i /tmp/b29_c1.ncu-rep --page raw --csv > gpurun_out/b29_raw.csv 2>/dev/null
ncu -i /tmp/b29_c1.ncu-rep --page source --csv > gpurun_out/b29_source.csv 2>/dev/null
ncu -i /tmp/b29_c1.ncu-rep --page details > gpurun_out/b29_details.txt 2>/dev/null
ls -la /tmp/b29_c1.ncu-rep gpurun_out/ >> gpurun_out/b29_ncu.log
tail -5 gpurun_out/b29_ncu.log
