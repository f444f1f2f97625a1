#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/b11_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/b11_pytest.log
timeout 300 python tools/r2_probe.py --what c1 > gpurun_out/b11_probe_c1.log 2>&1
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/b11_bench_launches.csv python bench.py --steps 2 --warmup 3 --skip-extras > gpurun_out/b11_ncu_launches.log 2>&1
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"spmv_ell|spmv_vector8|proj_uniform|commit|panel_reduce" -c 64 -o gpurun_out/b11_c5 python bench.py --steps 1 --warmup 3 --skip-extras > gpurun_out/b11_ncu_c5.log 2>&1
tail -3 gpurun_out/b11_pytest.log; cut -c1-200 gpurun_out/b11_probe_c1.log
