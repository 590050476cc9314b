# per-kernel A/B of the variant libraries in nisqa_b200/exp (tools/tc_ab_build.sh) - no tests
TAG=${1:-rXX}
mkdir -p gpurun_out
for v in $AB_VARIANTS; do timeout 120 python tools/ab_kernels.py --lib nisqa_b200/exp/libnisqa_$v.so --tag $v 2>&1 | grep "^\[" | tee -a gpurun_out/${TAG}_ab_kernels.txt; done
timeout 120 python tools/ab_kernels.py --tag default 2>&1 | grep "^\[" | tee -a gpurun_out/${TAG}_ab_kernels.txt
