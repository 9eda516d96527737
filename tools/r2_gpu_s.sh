#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -x -q -k "fixed_k or stress or subpixel" > gpurun_out/s_pytest.log 2>&1; tail -2 gpurun_out/s_pytest.log
timeout 300 python tools/stress_bench.py 2>&1 | tail -1
F=20000 timeout 600 ncu --set full --import-source on --clock-control none -k regex:ransac_fixedk -c 1 -o gpurun_out/r2_prof_fixedk_v4 -f python tools/stress_bench.py > gpurun_out/s_ncu.log 2>&1
ncu -i gpurun_out/r2_prof_fixedk_v4.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h=rows[0]; v=rows[2]
for k in ('gpu__time_duration.sum','sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread'):
    print(k, v[h.index(k)] if k in h else 'n/a')
"
