#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/fixedk_variants.py > gpurun_out/z5_fixedk_variants.txt 2>&1; cat gpurun_out/z5_fixedk_variants.txt | cut -c1-420
