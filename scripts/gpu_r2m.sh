set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2m_test.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2m_test.log
tail -4 gpurun_out/r2m_test.log
DDLO_PASS0_OVERLAP=1 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_batch.py -m gpu -x -q -k "align or batch or sequence or protocol" > gpurun_out/r2m_test_overlap.log 2>&1; tail -2 gpurun_out/r2m_test_overlap.log
timeout 600 python bench.py --steps 200 --no-cpu-baseline > gpurun_out/r2m_bench.json 2> gpurun_out/r2m_bench.err
timeout 300 python bench.py --steps 20 --no-cpu-baseline --no-c1 --batched-streams 8 --batched-units 2048 > gpurun_out/r2m_bench_l8.json 2>/dev/null
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_write.sum,dram__bytes_read.sum,smsp__inst_executed_op_local_ld.sum,smsp__inst_executed_op_local_st.sum,smsp__inst_executed.sum --clock-control none -k regex:"k_align" --launch-skip 2 -c 1 --csv --log-file gpurun_out/r2m_align_metrics.csv python profiles/profile_step.py 3 > gpurun_out/r2m_ncu.log 2>&1
grep -v "^==" gpurun_out/r2m_align_metrics.csv | cut -d, -f13- 
