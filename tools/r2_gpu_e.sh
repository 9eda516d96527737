#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-propagated > gpurun_out/e_bench.json 2> gpurun_out/e_bench.err
tail -3 gpurun_out/e_bench.err
python -c "
import json; d=json.load(open('gpurun_out/e_bench.json')); print(json.dumps({k:d[k] for k in ('value','e2e','full_match')}, indent=1))"
timeout 900 tools/variants_check.sh > gpurun_out/e_variants.log 2>&1; tail -30 gpurun_out/e_variants.log
