#!/bin/bash
# compute-sanitizer passes over the keypoint-propagation kernels (run on the GPU box); logs to gpurun_out/
mkdir -p gpurun_out
T="tests/test_gpu_flow.py::test_gray_pyramid_matches_oracle tests/test_gpu_flow.py::test_filter_flow_matches_oracle tests/test_gpu_flow.py::test_merge_and_calibrate_match_oracle tests/test_gpu_flow.py::test_sparse_cadence_very_short_clips tests/test_gpu_flow.py::test_sparse_cadence_golden_from_reference tests/test_gpu_parity.py::test_subpixel_refinement_extension"
for tool in memcheck racecheck initcheck; do
  timeout 1200 compute-sanitizer --tool $tool --print-limit 20 python -m pytest $T -m gpu -q -x > gpurun_out/sanitizer_flow_$tool.log 2>&1
  echo "== $tool: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/sanitizer_flow_$tool.log | tail -1) | $(grep -E 'passed|failed' gpurun_out/sanitizer_flow_$tool.log | tail -1)"
done
