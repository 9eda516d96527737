#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/j_pytest.log 2>&1; tail -5 gpurun_out/j_pytest.log
timeout 300 python tools/cascade_bench.py > gpurun_out/j_cascade_bench.json 2>gpurun_out/j_cascade_bench.err; cat gpurun_out/j_cascade_bench.json
timeout 300 python tools/stress_bench.py > gpurun_out/j_stress.log 2>&1; cat gpurun_out/j_stress.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/j_launches_stress.csv python tools/stress_bench.py > /dev/null 2>&1
grep -E "fixedk|refit|cascade" gpurun_out/j_launches_stress.csv | awk -F'","' '{print $5, $NF}' | tail -4
