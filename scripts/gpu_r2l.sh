set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for i in 1 2; do
timeout 300 python profiles/overlap_ab.py > gpurun_out/r2l_on_$i.log 2>&1; tail -2 gpurun_out/r2l_on_$i.log
DDLO_NO_PASS0_OVERLAP=1 timeout 300 python profiles/overlap_ab.py > gpurun_out/r2l_off_$i.log 2>&1; tail -2 gpurun_out/r2l_off_$i.log
done
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2l_test.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2l_test.log
tail -4 gpurun_out/r2l_test.log
timeout 600 python benchmarks/c3_sequence.py --cpp > gpurun_out/r2l_c3_cpp.json 2> gpurun_out/r2l_c3_cpp.err
DDLO_NO_PASS0_OVERLAP=1 timeout 600 python benchmarks/c3_sequence.py --cpp > gpurun_out/r2l_c3_cpp_off.json 2> /dev/null
