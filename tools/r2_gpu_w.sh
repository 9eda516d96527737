#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_flow.py -m gpu -q -x > gpurun_out/w_pytest.log 2>&1; tail -2 gpurun_out/w_pytest.log
timeout 600 python tools/soak_parity.py > gpurun_out/w_soak.txt 2>&1; tail -1 gpurun_out/w_soak.txt | cut -c1-200
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:refit -c 30 --csv --log-file gpurun_out/w_launches_refit.csv python tools/flow_bench.py > /dev/null 2>&1
grep refit gpurun_out/w_launches_refit.csv | awk -F'","' '{s+=$NF; n++} END {print "refit_warp in flow bench: mean", s/n/1000, "us over", n}'
timeout 300 python tools/few_inlier_bench.py > gpurun_out/w_few_inlier.json 2>&1; cat gpurun_out/w_few_inlier.json
