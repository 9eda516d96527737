#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/flow_bench.py 2>&1 | tail -1 | cut -c1-330
timeout 900 python -m pytest tests/test_gpu_flow.py -m gpu -q -x > gpurun_out/y_pytest.log 2>&1; tail -2 gpurun_out/y_pytest.log
