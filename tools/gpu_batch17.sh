#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"proj_tile_kernel|pava_tile_rows_kernel" -c 2 --launch-skip 6 -o /tmp/b17_c3 python tools/c3_run.py > gpurun_out/b17_c3.log 2>&1
ncu -i /tmp/b17_c3.ncu-rep --page raw --csv > gpurun_out/b17_c3_raw.csv 2>/dev/null
ncu -i /tmp/b17_c3.ncu-rep --page source --csv > gpurun_out/b17_c3_source.csv 2>/dev/null
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"pava_small_kernel" -c 1 --launch-skip 2 -o /tmp/b17_k64 python tools/one_pava.py 64 1562500 > gpurun_out/b17_k64.log 2>&1
ncu -i /tmp/b17_k64.ncu-rep --page raw --csv > gpurun_out/b17_k64_raw.csv 2>/dev/null
ncu -i /tmp/b17_k64.ncu-rep --page source --csv > gpurun_out/b17_k64_source.csv 2>/dev/null
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"pava_small_kernel" -c 1 --launch-skip 2 -o /tmp/b17_k16 python tools/one_pava.py 16 6250000 > gpurun_out/b17_k16.log 2>&1
ncu -i /tmp/b17_k16.ncu-rep --page raw --csv > gpurun_out/b17_k16_raw.csv 2>/dev/null
ncu -i /tmp/b17_k16.ncu-rep --page source --csv > gpurun_out/b17_k16_source.csv 2>/dev/null
du -sh gpurun_out; tail -2 gpurun_out/b17_*.log
