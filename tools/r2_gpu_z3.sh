#!/bin/bash
# ncu --set full of the new K1, the bench line after the assembler change
mkdir -p gpurun_out
timeout 300 ncu --set full --import-source on --clock-control none -k regex:preprocess_kernel -s 2 -c 1 -o gpurun_out/r2_prof_preprocess_pair -f python tools/k1_bench.py --once > gpurun_out/z3_ncu.log 2>&1; tail -2 gpurun_out/z3_ncu.log
ncu -i gpurun_out/r2_prof_preprocess_pair.ncu-rep --page raw --csv > gpurun_out/r2_prof_preprocess_pair_raw.csv 2>/dev/null; ls -la gpurun_out/r2_prof_preprocess_pair*
timeout 400 python bench.py > gpurun_out/z3_bench.json 2> gpurun_out/z3_bench.err; python - <<'P'
import json
try:
    d = json.loads(open("gpurun_out/z3_bench.json").read().strip().splitlines()[-1])
    print("value", round(d["value"]), "ms", round(d["ms_per_step"], 3), "sustained", round(d["sustained"]["ms_per_step"], 3), d["kernel_ms"]["decode_K2"], d["kernel_ms"]["synth_fit_select_project"], d["kernel_ms"]["preprocess_K1"], "frac", d["roofline"]["frac"], d["clocks"])
    print("e2e", round(d["e2e"]["value"]), d["api_e2e"]["us_per_frame"], "full_match", d["full_match"]["value"], d["full_match"]["with_dict_on_rank0"], d["full_match"]["rank0_host_split_s"])
except Exception as e:
    print("bench parse failed", e); print(open("gpurun_out/z3_bench.err").read()[-1500:])
P
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/z3_bench_ref.json 2>/dev/null; cut -c1-400 gpurun_out/z3_bench_ref.json
