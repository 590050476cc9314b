#!/bin/bash
# One GPU-box call that produces everything a round's profiles/ entry needs.
#   gpurun --timeout 900 -- 'bash tools/profile_round.sh r01d'
# Outputs land in gpurun_out/<tag>_*; tools/ncu_summary.py turns the raw ncu pages into the
# tracked profiles/<tag>_* summaries.
TAG=${1:-rXX}
SKIP_TESTS=${2:-0}
OUT=gpurun_out
mkdir -p $OUT
BENCH_PROF="python bench.py --steps 2 --warmup 1 --warmup-seconds 0 --skip-cpu"
if [ "$SKIP_TESTS" != "1" ]; then
  timeout 600 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1
  echo "pytest exit $?"; tail -3 $OUT/${TAG}_pytest.log
fi
timeout 600 python bench.py > $OUT/${TAG}_bench_n1.json 2> $OUT/${TAG}_bench.err
echo "bench exit $?"; cut -c1-400 $OUT/${TAG}_bench_n1.json
# launch list of the same command (cold-cache, serialised): one warm step skipped, two steps listed
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 15 -c 30 --csv \
    --log-file $OUT/${TAG}_launches_bench_steps2.csv $BENCH_PROF > $OUT/${TAG}_ncu_launches.log 2>&1
echo "ncu launches exit $?"
# full capture of one step (15 launches), raw page exported on the box
timeout 600 ncu --set full --clock-control none --launch-skip 15 -c 15 -f -o $OUT/${TAG}_full \
    $BENCH_PROF > $OUT/${TAG}_ncu_full.log 2>&1
echo "ncu full exit $?"
ncu -i $OUT/${TAG}_full.ncu-rep --page raw --csv > $OUT/${TAG}_full_raw.csv 2>/dev/null
ls -la $OUT | head -30
