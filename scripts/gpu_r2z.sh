set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 900 python bench.py > gpurun_out/r2z_bench.json 2> gpurun_out/r2z_bench.err ) 2> gpurun_out/r2z_bench_time.txt
( time timeout 900 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r2z_ref.json 2> gpurun_out/r2z_ref.err ) 2> gpurun_out/r2z_ref_time.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2z_test.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2z_test.log
tail -3 gpurun_out/r2z_test.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2z_smoke.log 2>&1; tail -1 gpurun_out/r2z_smoke.log
timeout 300 python bench.py --steps 50 --no-cpu-baseline --no-c1 --no-extra-configs --batched-wave 256 --batched-units 4096 > gpurun_out/r2z_w256.json 2> gpurun_out/r2z_w256.err
timeout 600 python benchmarks/c5_batch.py --wave 128 --out gpurun_out/r2z_c5_w128.json > /dev/null 2> gpurun_out/r2z_c5.err
timeout 600 python benchmarks/c5_batch.py --out gpurun_out/r2z_c5_w64.json > /dev/null 2>> gpurun_out/r2z_c5.err
