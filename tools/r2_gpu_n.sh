#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/n_pytest.log 2>&1; tail -4 gpurun_out/n_pytest.log
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/n_bench.json 2> gpurun_out/n_bench.err; tail -3 gpurun_out/n_bench.err
python -c "
import json; d=json.load(open('gpurun_out/n_bench.json')); print(d['value'], d['e2e']['value']); print(json.dumps(d['api_e2e'],indent=1)[:1800])"
