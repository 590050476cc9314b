# parity tests of a variant library (NISQA_LIB) + its per-kernel times next to the default build
#   AB_VARIANT=dit bash tools/ab_variant_tests.sh TAG
TAG=${1:-rXX}
mkdir -p gpurun_out
NISQA_LIB=$PWD/nisqa_b200/exp/libnisqa_$AB_VARIANT.so timeout 500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q \
    -k "stage or golden or edge or max_length or full_size or filterbank" > gpurun_out/${TAG}_pytest_$AB_VARIANT.log 2>&1
echo "pytest($AB_VARIANT) exit $?"; tail -2 gpurun_out/${TAG}_pytest_$AB_VARIANT.log
timeout 120 python tools/ab_kernels.py --lib nisqa_b200/exp/libnisqa_$AB_VARIANT.so --tag $AB_VARIANT 2>&1 | grep "^\[" | tee -a gpurun_out/${TAG}_ab_kernels.txt
timeout 120 python tools/ab_kernels.py --tag default 2>&1 | grep "^\[" | tee -a gpurun_out/${TAG}_ab_kernels.txt
