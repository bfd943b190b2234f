set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_batch.py -m gpu -x -q > gpurun_out/r2x_test.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2x_test.log
tail -5 gpurun_out/r2x_test.log
bash scripts/gpu_variants.sh "$@" 2>&1 | grep VARIANT
