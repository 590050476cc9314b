# one GPU call: parity tests, per-kernel A/B of the variant libraries in nisqa_b200/exp (tools/tc_ab_build.sh), phase timing
TAG=${1:-rXX}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/${TAG}_pytest.log
for v in $AB_VARIANTS; do timeout 120 python tools/ab_kernels.py --lib nisqa_b200/exp/libnisqa_$v.so --tag $v 2>&1 | grep "^\[" | tee -a gpurun_out/${TAG}_ab_kernels.txt; done
timeout 120 python tools/ab_kernels.py --tag default 2>&1 | grep "^\[" | tee -a gpurun_out/${TAG}_ab_kernels.txt
timeout 200 python tools/pipe_timing.py 2>&1 | grep "^conv" | tee gpurun_out/${TAG}_phase_cycles.txt
