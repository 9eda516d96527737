#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/h_pytest.log 2>&1; tail -6 gpurun_out/h_pytest.log
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k cascade -s > gpurun_out/h_cascade.log 2>&1; grep -i "cascade:" gpurun_out/h_cascade.log
timeout 600 python tools/soak_parity.py > gpurun_out/h_soak.txt 2>&1; tail -2 gpurun_out/h_soak.txt
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/h_bench.json 2> gpurun_out/h_bench.err
python -c "
import json; d=json.load(open('gpurun_out/h_bench.json')); print(d['value'], d['kernel_ms'], d['propagated_cadence']['ms_per_clip'], d['e2e']['value'])"
