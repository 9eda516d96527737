#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_flow.py -m gpu -q -s -k "carve_out or thread_and_warp_kernels or sharded" > gpurun_out/c_pytest.log 2>&1; tail -4 gpurun_out/c_pytest.log; grep -h "inlier fits" gpurun_out/c_pytest.log
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/c_bench_2gpu.json 2> gpurun_out/c_bench_2gpu.err
tail -3 gpurun_out/c_bench_2gpu.err | cut -c1-300
python -c "
import json; d=json.load(open('gpurun_out/c_bench_2gpu.json')); print(json.dumps({k:d[k] for k in ('value','sustained','e2e','api_e2e','h2d_probe','full_match')}, indent=1))"
