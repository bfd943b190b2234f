set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2g_test.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2g_test.log
tail -5 gpurun_out/r2g_test.log
timeout 600 python bench.py --steps 200 --no-cpu-baseline > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err
timeout 600 python benchmarks/c3_sequence.py --cpp > gpurun_out/r2g_c3_cpp.json 2> gpurun_out/r2g_c3_cpp.err
timeout 600 python benchmarks/c4_large.py > gpurun_out/r2g_c4.json 2> gpurun_out/r2g_c4.err
