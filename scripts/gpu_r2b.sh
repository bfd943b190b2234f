set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2b_test.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b_test.log
tail -5 gpurun_out/r2b_test.log
timeout 300 python bench.py --steps 100 --no-cpu-baseline --no-c1 > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err
for W in 16 32 64; do timeout 300 python benchmarks/c5_batch.py --pairs 1024 --lanes 8 --wave $W > gpurun_out/r2b_c5_w$W.json 2>&1; done
timeout 300 python benchmarks/c5_batch.py --pairs 1024 --lanes 4 --wave 32 > gpurun_out/r2b_c5_l4.json 2>&1
timeout 300 python benchmarks/c5_batch.py --pairs 1024 --lanes 8 --mode lanes > gpurun_out/r2b_c5_lanes.json 2>&1
for W in 16 64; do timeout 300 python bench.py --steps 20 --no-cpu-baseline --no-c1 --batched-streams 8 --batched-wave $W > gpurun_out/r2b_bench_w$W.json 2>&1; done
