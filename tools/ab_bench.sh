# whole-step A/B: bench.py (value / e2e, no CPU legs) for the default library and the variants in AB_VARIANTS, plus the
# front-end's frame-pairs-per-CTA sweep
TAG=${1:-rXX}
mkdir -p gpurun_out
for v in default $AB_VARIANTS; do
  if [ "$v" = default ]; then unset NISQA_LIB; else export NISQA_LIB=$PWD/nisqa_b200/exp/libnisqa_$v.so; fi
  timeout 200 python bench.py --skip-cpu --steps 150 > gpurun_out/${TAG}_bench_$v.json 2>/dev/null
  python - <<PY | tee -a gpurun_out/${TAG}_ab_bench.txt
import json
d=json.load(open("gpurun_out/${TAG}_bench_$v.json"))
print("[$v] value %.0f  e2e %.0f  ms/step %.4f  parity %.2e" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["parity_max_abs_vs_oracle"]))
PY
done
unset NISQA_LIB
timeout 120 python tools/fe_sweep.py 1 2 3 4 6 8 2>&1 | grep ppc | tee gpurun_out/${TAG}_fe_sweep.txt
