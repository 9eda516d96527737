#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_upload.py tests/test_gpu_parity.py -m gpu -q -x -k "upload or drop_in or golden_reference or sharded" > gpurun_out/o_pytest.log 2>&1; tail -3 gpurun_out/o_pytest.log
timeout 300 python tools/upload_probe.py 2>&1 | grep -E "egl_upload_frames|identical"
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/o_bench.json 2> gpurun_out/o_bench.err; tail -3 gpurun_out/o_bench.err
python -c "
import json; d=json.load(open('gpurun_out/o_bench.json')); print(d['value'], d['e2e']['value']); print(json.dumps(d['api_e2e']['us_per_frame']), d['api_e2e']['h2d_GBps_inside_upload_calls'], d['api_e2e']['frames_in_pinned_memory']['value'])"
