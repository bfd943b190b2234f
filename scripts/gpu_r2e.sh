set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2e_test.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2e_test.log
tail -5 gpurun_out/r2e_test.log
for L in 4 8; do timeout 300 python benchmarks/c5_batch.py --pairs 2048 --lanes $L --wave 32 > gpurun_out/r2e_c5_l${L}.json 2>&1; done
timeout 300 python __graft_entry__.py > gpurun_out/r2e_entry.log 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2e_smoke.log 2>&1
