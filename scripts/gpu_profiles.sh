# The ncu captures behind profiles/r02_*:  gpurun --timeout 2400 -- 'bash scripts/gpu_profiles.sh'
# then, here:  python profiles/ncu_summary.py gpurun_out/r02_prof_step.ncu-rep (and ..._batch) > profiles/r02_ncu_full_summary.txt,
#              python profiles/make_align_traffic.py gpurun_out/r02_prof_step.ncu-rep, python profiles/launch_summary.py <csv> [units]
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python profiles/profile_step.py 3 > gpurun_out/prof_profile_step.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_align|k_knn|k_morton_sort|k_emit|k_cells" --launch-skip 10 -c 5 -f -o gpurun_out/r02_prof_step python profiles/profile_step.py 3 > gpurun_out/prof_ncu_step.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_batch_search|k_batch_lin|k_batch_err" --launch-skip 3 -c 3 -f -o gpurun_out/r02_prof_batch python profiles/batch_profile.py 64 32 4 c2 > gpurun_out/prof_ncu_batch.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-c1 --no-extra-configs --batched-streams 0 > gpurun_out/prof_ncu_launches.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -c 8000 --csv --log-file gpurun_out/r02_batch_launches.csv python profiles/batch_profile.py 64 32 4 c2 > gpurun_out/prof_ncu_batch_launches.log 2>&1
ls -la gpurun_out/*.ncu-rep
