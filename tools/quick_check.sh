TAG=$1
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/${TAG}_pytest.log
timeout 300 python bench.py --skip-cpu > gpurun_out/${TAG}_bench_quick.json 2> gpurun_out/${TAG}_bench.err; echo "bench exit $?"
python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench_quick.json"))
print("value",d["value"],"e2e",d["e2e"]["value"],"ms",d["ms_per_step"],"parity",d["parity_max_abs_vs_oracle"])
print(d["kernel_ms_per_step"])
PY
