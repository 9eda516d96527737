#!/bin/bash
# upload ring: slice size sweep (variants build), then refit variant at 50 k frames, then restore the shipped build
mkdir -p gpurun_out
EGL_BENCH_VARIANTS=1 python -m eagle_b200.build --force > /dev/null 2>&1
for kb in 256 512 2048 4096; do echo "== slice $kb KB"; EGL_UPLOAD_SLICE_KB=$kb timeout 200 python tools/upload_probe.py 2>&1 | grep -E "egl_upload_frames +(4|8|16) threads"; done
echo "== refit variants at F=50000"
EGL_REFIT_VARIANT=1 timeout 200 python tools/stress_bench.py 2>&1 | tail -1
EGL_REFIT_VARIANT=2 timeout 200 python tools/stress_bench.py 2>&1 | tail -1
python -m eagle_b200.build --force > /dev/null 2>&1
