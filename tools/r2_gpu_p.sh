#!/bin/bash
# round-2 record run: the bench line as the driver runs it (all extras), the reference arm, the launch list of the same command
set -x
mkdir -p gpurun_out
timeout 1500 python bench.py > gpurun_out/p_bench.json 2> gpurun_out/p_bench.err; tail -3 gpurun_out/p_bench.err
timeout 900 python bench.py --impl reference > gpurun_out/p_bench_reference.json 2> gpurun_out/p_bench_reference.err; tail -2 gpurun_out/p_bench_reference.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/p_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > /dev/null 2>&1
python -c "
import json; d=json.load(open('gpurun_out/p_bench.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['cpu_baseline']); print(d['ransac_stress']['value'], d['propagated_cadence']['ms_per_clip'], d['full_match']['value'])
r=json.load(open('gpurun_out/p_bench_reference.json')); print(r['value'], r.get('cpu_baseline'))"
