#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_flow.py -m gpu -q -x > gpurun_out/u_pytest.log 2>&1; tail -3 gpurun_out/u_pytest.log
timeout 600 python tools/soak_parity.py > gpurun_out/u_soak.txt 2>&1; tail -1 gpurun_out/u_soak.txt | cut -c1-300
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:refit -c 30 --csv --log-file gpurun_out/u_launches_refit.csv python tools/flow_bench.py > /dev/null 2>&1
grep refit gpurun_out/u_launches_refit.csv | awk -F'","' '{s+=$NF; n++} END {print "refit_warp in flow bench: mean", s/n/1000, "us over", n}'
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/u_bench.json 2> gpurun_out/u_bench.err
python -c "
import json; d=json.load(open('gpurun_out/u_bench.json')); print(d['value'], d['kernel_ms'], d['propagated_cadence']['ms_per_clip'], d['e2e']['value'])"
