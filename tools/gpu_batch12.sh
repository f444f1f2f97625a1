#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/b12_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/b12_pytest.log
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/b12_bench_launches.csv python bench.py --steps 2 --warmup 3 --skip-extras > gpurun_out/b12_ncu_launches.log 2>&1
timeout 900 ncu --profile-from-start off --set full --clock-control none -k regex:"spmv_ell|spmv_vector8|proj_uniform|commit|panel_reduce" -c 31 -o /tmp/b12_c5 python bench.py --steps 1 --warmup 3 --skip-extras > gpurun_out/b12_ncu_c5.log 2>&1
ncu -i /tmp/b12_c5.ncu-rep --page raw --csv > gpurun_out/b12_c5_raw.csv 2>/dev/null
ls -la /tmp/b12_c5.ncu-rep >> gpurun_out/b12_ncu_c5.log
du -sh gpurun_out >> gpurun_out/b12_ncu_c5.log
tail -3 gpurun_out/b12_pytest.log
