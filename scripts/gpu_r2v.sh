set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 900 python bench.py > gpurun_out/r2v_bench.json 2> gpurun_out/r2v_bench.err ) 2> gpurun_out/r2v_bench_time.txt
( time timeout 900 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r2v_ref.json 2> gpurun_out/r2v_ref.err ) 2> gpurun_out/r2v_ref_time.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2v_test.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2v_test.log
tail -3 gpurun_out/r2v_test.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2v_smoke.log 2>&1; tail -1 gpurun_out/r2v_smoke.log
