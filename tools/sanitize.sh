#!/bin/bash
# compute-sanitizer passes over small-shape GPU tests (run on the GPU box); logs to gpurun_out/
mkdir -p gpurun_out
T="tests/test_gpu_parity.py::test_decode_edge_cases_golden tests/test_gpu_parity.py::test_decode_special_values tests/test_gpu_parity.py::test_find_homography_golden_cases tests/test_gpu_parity.py::test_empty_and_degenerate_frames tests/test_gpu_parity.py::test_homography_cadence_and_failures tests/test_gpu_parity.py::test_dict_equals_golden_reference_run tests/test_gpu_parity.py::test_fixed_k_bit_exact_against_c_mirror tests/test_gpu_parity.py::test_preprocess_golden_checksums"
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python -m pytest $T -m gpu -q -x > gpurun_out/sanitizer_$tool.log 2>&1
  echo "== $tool: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/sanitizer_$tool.log | tail -1) | $(grep -E 'passed|failed' gpurun_out/sanitizer_$tool.log | tail -1)"
done
# TMA-ring decode variants and the thread-per-frame refit under memcheck as well
for v in 0 2 6; do
  EGL_DECODE_VARIANT=$v timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_parity.py::test_decode_edge_cases_golden tests/test_gpu_parity.py::test_decode_special_values -m gpu -q > gpurun_out/sanitizer_memcheck_decode_v$v.log 2>&1
  echo "== memcheck decode variant $v: $(grep -E 'ERROR SUMMARY' gpurun_out/sanitizer_memcheck_decode_v$v.log | tail -1) | $(grep -E 'passed|failed' gpurun_out/sanitizer_memcheck_decode_v$v.log | tail -1)"
done
