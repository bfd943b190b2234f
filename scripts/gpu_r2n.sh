set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2n_test.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2n_test.log
tail -4 gpurun_out/r2n_test.log
timeout 600 python bench.py --steps 200 --no-cpu-baseline > gpurun_out/r2n_bench.json 2> gpurun_out/r2n_bench.err
timeout 300 python bench.py --steps 20 --no-cpu-baseline --no-c1 --batched-streams 8 --batched-units 2048 > gpurun_out/r2n_bench_l8.json 2>/dev/null
timeout 600 python benchmarks/c3_sequence.py --cpp > gpurun_out/r2n_c3_cpp.json 2> gpurun_out/r2n_c3_cpp.err
