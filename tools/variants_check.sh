#!/bin/bash
# A/B build: rebuild the library WITH the measurement variants (-DEGL_BENCH_VARIANTS), run the cross-check tests that
# need them and the decode / preprocess sweeps, then restore the shipped build (one kernel per job).
set -x
mkdir -p gpurun_out
EGL_BENCH_VARIANTS=1 python -m eagle_b200.build --force > /dev/null
python -c "from eagle_b200 import _native; assert _native.lib.egl_build_flags() == 1"
timeout 900 python -m pytest tests/test_gpu_flow.py -m gpu -q -k "warp_and_thread" > gpurun_out/variants_pytest.log 2>&1; tail -3 gpurun_out/variants_pytest.log
timeout 900 python tools/sweep.py > gpurun_out/variants_sweep.log 2>&1; tail -20 gpurun_out/variants_sweep.log
python -m eagle_b200.build --force > /dev/null
python -c "from eagle_b200 import _native; assert _native.lib.egl_build_flags() == 0"
