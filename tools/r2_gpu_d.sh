#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/d_pytest.log 2>&1
tail -12 gpurun_out/d_pytest.log
grep -h "fixed-K vs cv2\|inlier fits" gpurun_out/d_pytest.log
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k "cv2_arithmetic or carve_out" > gpurun_out/d_pytest_report.log 2>&1
grep -h "fixed-K vs cv2\|inlier fits" gpurun_out/d_pytest_report.log
