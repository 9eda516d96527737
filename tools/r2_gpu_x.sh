#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_flow.py -m gpu -q -x > gpurun_out/x_pytest.log 2>&1; tail -2 gpurun_out/x_pytest.log
timeout 300 python tools/flow_bench.py 2>&1 | tail -1 | cut -c1-400
timeout 700 python tools/soak_flow.py > gpurun_out/x_soak_flow.txt 2>&1; tail -1 gpurun_out/x_soak_flow.txt
