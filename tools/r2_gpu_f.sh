#!/bin/bash
set -x
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/f_topo.txt 2>&1; lscpu | grep -E "^CPU\(s\)|Model name|NUMA" >> gpurun_out/f_topo.txt; free -g >> gpurun_out/f_topo.txt
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/f_bench_8gpu.json 2> gpurun_out/f_bench_8gpu.err
tail -3 gpurun_out/f_bench_8gpu.err | cut -c1-300
python -c "
import json; d=json.load(open('gpurun_out/f_bench_8gpu.json')); print(json.dumps({k:d[k] for k in ('value','sustained','e2e','api_e2e','h2d_probe','full_match')}, indent=1))"
