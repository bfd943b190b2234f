set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fuzz" > gpurun_out/r2s_test.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2s_test.log
tail -30 gpurun_out/r2s_test.log
