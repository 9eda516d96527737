#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q -x > gpurun_out/g_pytest.log 2>&1; tail -4 gpurun_out/g_pytest.log
timeout 600 python tools/soak_parity.py > gpurun_out/g_soak.txt 2>&1; tail -4 gpurun_out/g_soak.txt
timeout 600 python tools/soak_flow.py > gpurun_out/g_soak_flow.txt 2>&1; tail -3 gpurun_out/g_soak_flow.txt
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/g_bench.json 2> gpurun_out/g_bench.err
python -c "
import json; d=json.load(open('gpurun_out/g_bench.json')); print(d['value'], d['kernel_ms'], d['propagated_cadence']['ms_per_clip'])"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:refit -c 30 --csv --log-file gpurun_out/g_launches_refit.csv python tools/flow_bench.py > /dev/null 2>&1
grep -c refit gpurun_out/g_launches_refit.csv
