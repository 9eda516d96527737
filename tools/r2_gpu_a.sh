#!/bin/bash
# round-2 GPU call: fixed-K packed kernel -- parity, throughput, ncu; both occupancy caps (rebuilds fit.cu on the box)
set -x
mkdir -p gpurun_out
for cap in 7 6; do
  sed -i "s/__launch_bounds__(kFixedThreads, [0-9])/__launch_bounds__(kFixedThreads, $cap)/" eagle_b200/csrc/fit.cu
  python -m eagle_b200.build > /dev/null 2>&1
  timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fixed_k or stress or soak or subpixel" > gpurun_out/a_pytest_$cap.log 2>&1
  tail -2 gpurun_out/a_pytest_$cap.log
  timeout 300 python tools/stress_bench.py > gpurun_out/a_stress_$cap.log 2>&1
  cat gpurun_out/a_stress_$cap.log
  F=20000 timeout 600 ncu --set full --import-source on --clock-control none -k regex:ransac_fixedk -c 1 -o gpurun_out/r2_prof_fixedk_cap$cap -f python tools/stress_bench.py > gpurun_out/a_ncu_$cap.log 2>&1
done
