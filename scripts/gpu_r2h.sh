set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2h_test.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2h_test.log
tail -5 gpurun_out/r2h_test.log
timeout 600 python benchmarks/c3_sequence.py --cpp > gpurun_out/r2h_c3_cpp.json 2> gpurun_out/r2h_c3_cpp.err
timeout 600 python benchmarks/c4_large.py > gpurun_out/r2h_c4.json 2> gpurun_out/r2h_c4.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2h_c4_launches.csv python benchmarks/c4_large.py --steps 1 > gpurun_out/r2h_c4_ncu.log 2>&1
