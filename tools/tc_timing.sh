#!/bin/bash
set -e
cd "$(dirname "$0")/.."
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -fvisibility=hidden --shared \
  -DNISQA_TC_TIMING nisqa_b200/csrc/{engine,frontend,cnn,conv_tc,td}.cu -o nisqa_b200/libnisqa_b200.so -ldl
python tools/tc_timing.py
