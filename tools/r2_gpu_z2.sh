#!/bin/bash
# K1 word-load + dp4a path: A/B + parity, the whole GPU suite, the bench line, ncu capture of K1 and the launch list
mkdir -p gpurun_out
timeout 300 python tools/k1_bench.py > gpurun_out/z2_k1_ab.txt 2>&1; cat gpurun_out/z2_k1_ab.txt
timeout 120 python __graft_entry__.py --smoke 2>&1 | tail -2
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/z2_pytest.log 2>&1; tail -3 gpurun_out/z2_pytest.log
timeout 400 python bench.py > gpurun_out/z2_bench.json 2> gpurun_out/z2_bench.err; python - <<'P'
import json
try:
    d = json.loads(open("gpurun_out/z2_bench.json").read().strip().splitlines()[-1])
    print("value", round(d["value"]), "ms", round(d["ms_per_step"], 3), "sustained", round(d["sustained"]["ms_per_step"], 3), d["kernel_ms"]["decode_K2"], d["kernel_ms"]["synth_fit_select_project"], d["kernel_ms"]["preprocess_K1"], "frac", d["roofline"]["frac"], d["clocks"])
    print("e2e", round(d["e2e"]["value"]), "full_match", d["full_match"]["value"], d["full_match"]["with_dict_on_rank0"]["value"], "prop", d["propagated_cadence"]["ms_per_clip"], "stress", d["ransac_stress"]["value"])
except Exception as e:
    print("bench parse failed", e); print(open("gpurun_out/z2_bench.err").read()[-1500:])
P
timeout 300 ncu --set full --import-source on --clock-control none -k regex:preprocess_kernel -s 3 -c 1 -o gpurun_out/r2_prof_preprocess_pair -f python tools/k1_bench.py --once > gpurun_out/z2_ncu.log 2>&1; tail -2 gpurun_out/z2_ncu.log
ncu -i gpurun_out/r2_prof_preprocess_pair.ncu-rep --page raw --csv > gpurun_out/r2_prof_preprocess_pair_raw.csv 2>/dev/null; ls -la gpurun_out/r2_prof_preprocess_pair*
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/z2_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras --no-propagated > /dev/null 2>&1; wc -l gpurun_out/z2_launches_bench.csv
