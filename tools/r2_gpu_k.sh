#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "cascade or rejection or golden_cases or fixed_k or stress" > gpurun_out/k_pytest.log 2>&1; tail -3 gpurun_out/k_pytest.log
timeout 300 python tools/cascade_bench.py > gpurun_out/k_cascade_bench.json 2>gpurun_out/k_cascade_bench.err; cat gpurun_out/k_cascade_bench.json
timeout 300 python tools/stress_bench.py > gpurun_out/k_stress.log 2>&1; cat gpurun_out/k_stress.log
