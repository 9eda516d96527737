#!/bin/bash
# fixed-K sign-bit count: A/B under ncu + digests, the whole GPU suite on the shipped build, the fixed-K tests on the
# compare-and-add build as well, compute-sanitizer over the new kernels, the bench line, ncu --set full of the hypothesis kernel
mkdir -p gpurun_out
timeout 600 python tools/fixedk_variants.py > gpurun_out/z6_fixedk_variants.txt 2>&1; cut -c1-400 gpurun_out/z6_fixedk_variants.txt
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/z6_pytest.log 2>&1; tail -3 gpurun_out/z6_pytest.log
EAGLE_B200_LIBRARY=$PWD/tools/_variants/libeagle_b200_variants.so EGL_FIXEDK_VARIANT=10 timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q -k "fixed_k or stress" > gpurun_out/z6_pytest_v10.log 2>&1; tail -2 gpurun_out/z6_pytest_v10.log
timeout 300 compute-sanitizer --tool memcheck --print-limit 10 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "letterbox or views_and_strides or fixed_k_bit_exact" > gpurun_out/z6_sanitizer.log 2>&1; echo "memcheck: $(grep -E 'ERROR SUMMARY' gpurun_out/z6_sanitizer.log | tail -1) | $(grep -E 'passed|failed' gpurun_out/z6_sanitizer.log | tail -1)"
timeout 400 python bench.py > gpurun_out/z6_bench.json 2> gpurun_out/z6_bench.err; python - <<'P'
import json
try:
    d = json.loads(open("gpurun_out/z6_bench.json").read().strip().splitlines()[-1])
    print("value", round(d["value"]), "ms", round(d["ms_per_step"], 3), "sustained", round(d["sustained"]["ms_per_step"], 3), "frac", d["roofline"]["frac"], "traffic", d["roofline"]["traffic"], "e2e", round(d["e2e"]["value"]))
    print("stress", d["ransac_stress"]["value"], d["ransac_stress"]["ms_per_batch"], "full_match", d["full_match"]["value"], d["full_match"]["with_dict_on_rank0"]["value"], "prop", d["propagated_cadence"]["ms_per_clip"])
except Exception as e:
    print("bench parse failed", e); print(open("gpurun_out/z6_bench.err").read()[-1500:])
P
F=20000 timeout 300 ncu --set full --import-source on --clock-control none -k regex:ransac_fixedk -s 1 -c 1 -o gpurun_out/r2_prof_fixedk_v5 -f python tools/stress_bench.py > gpurun_out/z6_ncu.log 2>&1; tail -1 gpurun_out/z6_ncu.log
ncu -i gpurun_out/r2_prof_fixedk_v5.ncu-rep --page raw --csv > gpurun_out/r2_prof_ransac_fixedk_v5_raw.csv 2>/dev/null
ncu -i gpurun_out/r2_prof_fixedk_v5.ncu-rep --page source --csv > gpurun_out/r2_prof_ransac_fixedk_v5_source.csv 2>/dev/null; ls -la gpurun_out/r2_prof_*v5*
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/z6_launches_stress.csv python tools/stress_bench.py > gpurun_out/z6_stress.log 2>&1; tail -1 gpurun_out/z6_stress.log | cut -c1-300
