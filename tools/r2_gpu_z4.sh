#!/bin/bash
# 2 GPUs: run_sharded over NCCL against the reference's golden dict; the bench line at N=2 as the driver launches it
mkdir -p gpurun_out
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tools/sharded_check.py 2>&1 | tail -2
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 tools/sharded_flow_check.py 2>&1 | tail -2
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/z4_bench_2gpu.json 2> gpurun_out/z4_bench_2gpu.err; python - <<'P'
import json
try:
    d = json.loads(open("gpurun_out/z4_bench_2gpu.json").read().strip().splitlines()[-1])
    print("N=2 value", round(d["value"]), "ms", round(d["ms_per_step"], 3), "sustained", round(d["sustained"]["ms_per_step"], 3), d["kernel_ms"], "e2e", round(d["e2e"]["value"]))
    print("full_match", d["full_match"]["value"], d["full_match"]["with_dict_on_rank0"]["value"], "prop", d["propagated_cadence"]["ms_per_clip"], d["h2d_probe"]["per_rank_GBps"], d["h2d_probe"]["pageable_upload_per_rank_GBps"])
except Exception as e:
    print("bench parse failed", e); print(open("gpurun_out/z4_bench_2gpu.err").read()[-2000:])
P
