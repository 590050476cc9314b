#!/bin/bash
# GPU-box A/B driver; variant builds are prebuilt by tools/tc_ab_build.sh (nisqa_b200/exp/*.so travel with gpurun).
#   bash tools/tc_ab_build.sh            (here)        gpurun --timeout 400 -- 'bash tools/tc_ab.sh'
cd "$(dirname "$0")/.."
run() { timeout "$1" python tools/tc_ab.py "${@:2}" > /tmp/ab.log 2>&1; rc=$?; grep -v Warning /tmp/ab.log | tail -14; echo "rc=$rc"; return $rc; }
python -c "import torch" 2>/dev/null     # page the image in outside the timeouts
run 200 --split 1 --tag planes || { echo "default variant failed - stopping"; exit 0; }
run 150 --split 1 --wide 0x78 --tag wide
run 150 --split 0 --skip-check --tag tc_f32
[ -f nisqa_b200/exp/libnisqa_timing.so ] && run 150 --lib nisqa_b200/exp/libnisqa_timing.so --timing --skip-check --split 1 --tag T_planes
