set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2d_test.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2d_test.log
tail -5 gpurun_out/r2d_test.log
timeout 600 python bench.py --steps 200 > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err
for L in 4 8; do for W in 32 64; do timeout 300 python benchmarks/c5_batch.py --pairs 2048 --lanes $L --wave $W > gpurun_out/r2d_c5_l${L}_w$W.json 2>&1; done; done
for L in 4 8 16; do timeout 300 python bench.py --steps 20 --no-cpu-baseline --no-c1 --batched-streams $L --batched-wave 32 > gpurun_out/r2d_bench_l$L.json 2>&1; done
timeout 300 python bench.py --steps 20 --no-cpu-baseline --no-c1 --batched-streams 8 --batched-wave 64 --batched-units 2048 > gpurun_out/r2d_bench_l8w64.json 2>&1
timeout 300 python profiles/batch_profile.py 64 32 4 c2 > gpurun_out/r2d_bp_plain.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -c 8000 --csv --log-file gpurun_out/r2d_batch_launches.csv python profiles/batch_profile.py 64 32 4 c2 > gpurun_out/r2d_bp_ncu.log 2>&1
