#!/bin/bash
# fixed-K kernel with the prepare kernel split off: parity, throughput, ncu
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -x -q -k "fixed_k or stress or subpixel or cascade" > gpurun_out/i_pytest.log 2>&1
tail -3 gpurun_out/i_pytest.log
timeout 300 python tools/stress_bench.py > gpurun_out/i_stress.log 2>&1; cat gpurun_out/i_stress.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/i_launches_stress.csv python tools/stress_bench.py > /dev/null 2>&1
grep -E "fixedk|refit|cascade" gpurun_out/i_launches_stress.csv | awk -F'","' '{print $5, $NF}' | tail -8
F=20000 timeout 600 ncu --set full --import-source on --clock-control none -k regex:ransac_fixedk -c 1 -o gpurun_out/r2_prof_fixedk_v3 -f python tools/stress_bench.py > gpurun_out/i_ncu.log 2>&1
ncu -i gpurun_out/r2_prof_fixedk_v3.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h=rows[0]; v=rows[2]
for k in ('gpu__time_duration.sum','sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active','sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed','sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio','launch__registers_per_thread'):
    print(k, v[h.index(k)] if k in h else 'n/a')
"
