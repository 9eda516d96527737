#!/bin/bash
# the shipped build after the fixed-K default went back to compare + predicated add: fixed-K tests, bench line, stress launch list
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q -k "fixed_k or stress or preprocess" > gpurun_out/z7_pytest.log 2>&1; tail -2 gpurun_out/z7_pytest.log
timeout 400 python bench.py > gpurun_out/z7_bench.json 2> gpurun_out/z7_bench.err; python - <<'P'
import json
try:
    d = json.loads(open("gpurun_out/z7_bench.json").read().strip().splitlines()[-1])
    print("value", round(d["value"]), "ms", round(d["ms_per_step"], 3), "sustained", round(d["sustained"]["ms_per_step"], 3), "frac", d["roofline"]["frac"], "traffic", d["roofline"]["traffic"], "e2e", round(d["e2e"]["value"]))
    print("stress", d["ransac_stress"]["value"], d["ransac_stress"]["ms_per_batch"], "full_match", d["full_match"]["value"], d["full_match"]["with_dict_on_rank0"]["value"], "prop", d["propagated_cadence"]["ms_per_clip"])
except Exception as e:
    print("bench parse failed", e); print(open("gpurun_out/z7_bench.err").read()[-1500:])
P
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/z7_launches_stress.csv python tools/stress_bench.py > gpurun_out/z7_stress.log 2>&1; tail -1 gpurun_out/z7_stress.log | cut -c1-200
