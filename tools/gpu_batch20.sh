#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"proj_tile_kernel|proj_mid_kernel|proj_large_kernel" -c 3 --launch-skip 6 -o /tmp/b20_c3 python tools/c3_run.py > gpurun_out/b20_c3.log 2>&1
ncu -i /tmp/b20_c3.ncu-rep --page raw --csv > gpurun_out/b20_c3_raw.csv 2>/dev/null
ncu -i /tmp/b20_c3.ncu-rep --page source --csv > gpurun_out/b20_c3_source.csv 2>/dev/null
du -sh gpurun_out; tail -n 2 gpurun_out/b20_c3.log
