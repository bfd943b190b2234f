set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_batch.py -m gpu -x -q > gpurun_out/r2c_test.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c_test.log
tail -5 gpurun_out/r2c_test.log
DDLO_BATCH_TRACE=1 timeout 300 python benchmarks/c5_batch.py --pairs 512 --lanes 8 --wave 32 > gpurun_out/r2c_c5_trace.json 2> gpurun_out/r2c_c5_trace.err
DDLO_BATCH_TRACE=1 DDLO_BATCH_SKIP_ALIGN=1 timeout 300 python benchmarks/c5_batch.py --pairs 512 --lanes 8 --wave 32 > gpurun_out/r2c_c5_skip.json 2> gpurun_out/r2c_c5_skip.err
DDLO_BATCH_TRACE=1 timeout 300 python bench.py --steps 20 --no-cpu-baseline --no-c1 --batched-streams 8 --batched-wave 32 > gpurun_out/r2c_bench_trace.json 2> gpurun_out/r2c_bench_trace.err
DDLO_BATCH_TRACE=1 DDLO_BATCH_SKIP_ALIGN=1 timeout 300 python bench.py --steps 20 --no-cpu-baseline --no-c1 --batched-streams 8 --batched-wave 32 > gpurun_out/r2c_bench_skip.json 2> gpurun_out/r2c_bench_skip.err
DDLO_BATCH_SKIP_ALIGN=1 timeout 300 python bench.py --steps 20 --no-cpu-baseline --no-c1 --batched-streams 16 --batched-wave 32 > gpurun_out/r2c_bench_skip16.json 2> /dev/null
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2c_testall.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c_testall.log
tail -5 gpurun_out/r2c_testall.log
