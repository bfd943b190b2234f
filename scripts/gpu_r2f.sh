set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2f_test.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2f_test.log
tail -5 gpurun_out/r2f_test.log
timeout 600 python benchmarks/c3_sequence.py --cpp > gpurun_out/r2f_c3_cpp.json 2> gpurun_out/r2f_c3_cpp.err
timeout 600 python benchmarks/c3_sequence.py > gpurun_out/r2f_c3_py.json 2> gpurun_out/r2f_c3_py.err
timeout 600 python benchmarks/c3_sequence.py --cpp --voxel 0.25 > gpurun_out/r2f_c3_cpp_voxel.json 2> /dev/null
