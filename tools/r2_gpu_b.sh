#!/bin/bash
# round-2 GPU call B: whole GPU suite, smoke, 1-GPU bench
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/b_pytest.log 2>&1
tail -15 gpurun_out/b_pytest.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/b_smoke.log 2>&1
tail -3 gpurun_out/b_smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/b_bench.json 2> gpurun_out/b_bench.err
tail -5 gpurun_out/b_bench.err
python -c "
import json; d=json.load(open('gpurun_out/b_bench.json')); print(json.dumps({k:d[k] for k in ('value','e2e','api_e2e','h2d_probe','full_match')}, indent=1))"
