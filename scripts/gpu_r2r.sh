set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2r_test.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2r_test.log
tail -3 gpurun_out/r2r_test.log
for i in 1 2; do timeout 600 python benchmarks/c3_sequence.py --cpp > gpurun_out/r2r_c3_cpp_$i.json 2> gpurun_out/r2r_c3_cpp.err; done
timeout 600 python benchmarks/c3_sequence.py --cpp --voxel 0.25 > gpurun_out/r2r_c3_cpp_voxel.json 2> /dev/null
