#!/bin/bash
set -x
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/c_topo.txt 2>&1; lscpu | head -25 >> gpurun_out/c_topo.txt; free -g >> gpurun_out/c_topo.txt
timeout 300 python tools/host_register_probe.py > gpurun_out/c_register.log 2>&1
cat gpurun_out/c_register.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/c_bench_2gpu.json 2> gpurun_out/c_bench_2gpu.err
tail -5 gpurun_out/c_bench_2gpu.err
python -c "
import json; d=json.load(open('gpurun_out/c_bench_2gpu.json')); print(json.dumps({k:d[k] for k in ('value','e2e','api_e2e','h2d_probe','full_match')}, indent=1))"
timeout 600 python -m pytest tests/test_gpu_flow.py -m gpu -q -k "sharded" > gpurun_out/c_pytest.log 2>&1; tail -3 gpurun_out/c_pytest.log
