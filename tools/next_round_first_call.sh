#!/bin/bash
# Everything that was staged without GPU time at the end of round 1, in ONE GPU-box call (about 4 minutes):
#   bash tools/tc_ab_build.sh                                   (here, prebuilds nisqa_b200/exp/libnisqa_timing.so)
#   gpurun --timeout 420 -- 'bash tools/next_round_first_call.sh'
# 1. tcgen05.mma cost table (N, accumulator chains, co-resident CTAs)      -> gpurun_out/mma_probe.txt
# 2. gated GPU tests: conv_wide candidate kernels, ms_sr end-to-end ingest   -> gpurun_out/experimental_pytest.log
# 3. A/B: default plane pipeline vs conv_wide on conv3..6 (+ phase stamps)   -> gpurun_out/ab_wide.log
# 4. compute-sanitizer memcheck of the default path                          -> gpurun_out/<tag>_memcheck*.log
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import torch" 2>/dev/null
timeout 100 bash tools/mma_probe.sh > /dev/null 2>&1; echo "mma_probe exit $?"; head -40 gpurun_out/mma_probe.txt
NISQA_EXPERIMENTAL=1 timeout 150 python -m pytest tests -m gpu -q -k "conv_wide or ms_sr" > gpurun_out/experimental_pytest.log 2>&1
echo "experimental pytest exit $?"; tail -5 gpurun_out/experimental_pytest.log
{ timeout 120 python tools/tc_ab.py --split 1 --skip-check --tag planes;
  timeout 120 python tools/tc_ab.py --split 1 --wide 0x78 --tag wide;
  [ -f nisqa_b200/exp/libnisqa_timing.so ] && timeout 120 python tools/tc_ab.py --lib nisqa_b200/exp/libnisqa_timing.so --timing --skip-check --split 1 --tag T_planes; } 2>&1 | grep -v Warning > gpurun_out/ab_wide.log
cat gpurun_out/ab_wide.log
timeout 200 bash tools/sanitize.sh r02 2>&1 | tail -8
