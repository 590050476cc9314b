#!/bin/bash
# GPU-box experiment: weight-ring depth of the tcgen05 conv kernels
set -e
cd "$(dirname "$0")/.."
for cfg in "2 2 4" "9 9 9" "4 4 9" "3 9 6"; do
  set -- $cfg
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -fvisibility=hidden --shared \
    -DNISQA_TC_NS4=$1 -DNISQA_TC_NS5=$2 -DNISQA_TC_NS3=$3 nisqa_b200/csrc/{engine,frontend,cnn,conv_tc,td}.cu -o nisqa_b200/libnisqa_b200.so -ldl
  echo "== NS4=$1 NS5=$2 NS3=$3"
  python bench.py --steps 30 --skip-cpu 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
k = d['kernel_ms_per_step']
print('value %.0f' % d['value'], {x: round(k[x], 4) for x in ('conv2','conv3','conv4','conv5','conv6')})"
done
