#!/bin/bash
# GPU-box A/B driver; the variant builds are prebuilt by tools/tc_ab_build.sh (nisqa_b200/exp/*.so)
cd "$(dirname "$0")/.."
run() { timeout "$1" python tools/tc_ab.py "${@:2}" > /tmp/ab.log 2>&1; rc=$?; grep -v Warning /tmp/ab.log | tail -14; echo "rc=$rc"; return $rc; }
python -c "import torch" 2>/dev/null     # page the image in outside the timeouts
run 200 --split 1 --skip-check --sweep 0,50,100,150,200,0 --tag sweep
run 150 --lib nisqa_b200/exp/libnisqa_timing.so --timing --skip-check --split 1 --stagger 100 --tag T_stagger100
