#!/bin/bash
# round-2 GPU call B: whole GPU suite, smoke, 1-GPU bench, launch list of the stress fit
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/b_pytest.log 2>&1
tail -15 gpurun_out/b_pytest.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/b_smoke.log 2>&1
tail -3 gpurun_out/b_smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/b_bench.json 2> gpurun_out/b_bench.err
tail -5 gpurun_out/b_bench.err
cat gpurun_out/b_bench.json | head -c 6000
F=50000 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/b_launches_stress.csv python tools/stress_bench.py > /dev/null 2>&1
grep -c . gpurun_out/b_launches_stress.csv
