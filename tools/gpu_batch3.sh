#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/b3_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/b3_pytest.log
timeout 300 python tools/r2_probe.py --what c1 > gpurun_out/b3_probe_c1.log 2>&1
timeout 120 ./tools/probes/tma_gather_probe > gpurun_out/b3_tma_probe.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"spmv_ell|spmv_vector" --launch-skip 1 -c 6 -o gpurun_out/b3_spmv python tools/r2_probe.py --scale 0.125 --what spmv --panel-mb 48 > gpurun_out/b3_ncu.log 2>&1
tail -3 gpurun_out/b3_pytest.log; cat gpurun_out/b3_tma_probe.log
