#!/bin/bash
# gpurun --timeout 120 -- 'bash tools/mma_probe.sh'   (builds on the box: ~10 s, runs a few seconds)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I nisqa_b200/csrc tools/mma_probe.cu -o /tmp/mma_probe || exit 1
timeout 60 /tmp/mma_probe | tee gpurun_out/mma_probe.txt
