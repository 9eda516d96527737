#!/bin/bash
# step overlap (tail under K1), streamed run_sharded on a side stream, two uploads in flight, staging-copy variants
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_upload.py tests/test_gpu_fullsize.py "tests/test_gpu_parity.py" -m gpu -q -x -k "upload or streamed or sharded or full_clip or config0 or in_flight" > gpurun_out/z1_pytest.log 2>&1; tail -3 gpurun_out/z1_pytest.log
timeout 400 python bench.py > gpurun_out/z1_bench.json 2> gpurun_out/z1_bench.err; python - <<'P'
import json
try:
    d = json.loads(open("gpurun_out/z1_bench.json").read().strip().splitlines()[-1])
    print("value", round(d["value"]), "ms", round(d["ms_per_step"], 3), "sustained", round(d["sustained"]["ms_per_step"], 3), d["kernel_ms"]["decode_K2"], d["kernel_ms"]["synth_fit_select_project"], d["kernel_ms"]["preprocess_K1"])
    print("e2e", round(d["e2e"]["value"]), "upload GB/s in calls", d["api_e2e"]["h2d_GBps_inside_upload_calls"], "probe", d["h2d_probe"])
    print("full_match", d["full_match"]["value"], d["full_match"]["with_dict_on_rank0"]["value"], "prop", d["propagated_cadence"]["ms_per_clip"], "stress", d["ransac_stress"]["value"])
except Exception as e:
    print("bench parse failed", e); print(open("gpurun_out/z1_bench.err").read()[-1500:])
P
timeout 300 python tools/upload_variants_probe.py > gpurun_out/z1_upload_variants.txt 2>&1; cat gpurun_out/z1_upload_variants.txt
timeout 300 python tools/api_inflight_sweep.py > gpurun_out/z1_api_inflight.txt 2>&1; tail -8 gpurun_out/z1_api_inflight.txt
