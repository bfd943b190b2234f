set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version --format=csv > gpurun_out/r2a_smi.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_test.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_test.log
timeout 600 python bench.py --steps 200 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench rc=$?"
for L in 4 8; do timeout 300 python benchmarks/c5_batch.py --pairs 1024 --lanes $L --host-threads 2 > gpurun_out/r2a_c5_l$L.json 2> gpurun_out/r2a_c5_l$L.err; done
timeout 300 python benchmarks/c5_batch.py --pairs 1024 --lanes 8 --host-threads 4 > gpurun_out/r2a_c5_l8t4.json 2>&1
timeout 300 python benchmarks/c5_batch.py --pairs 1024 --lanes 8 --host-threads 1 > gpurun_out/r2a_c5_l8t1.json 2>&1
for S in 2 8; do timeout 300 python bench.py --steps 20 --no-cpu-baseline --no-c1 --batched-streams $S > gpurun_out/r2a_bench_s$S.json 2>&1; done
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python profiles/sanitize_case.py > gpurun_out/r2a_sanitizer_$tool.txt 2>&1; echo "$tool rc=$?" >> gpurun_out/r2a_sanitizer_$tool.txt
done
tail -3 gpurun_out/r2a_test.log
