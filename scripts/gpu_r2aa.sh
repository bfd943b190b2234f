set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python profiles/batch_profile.py 64 32 4 c2 > gpurun_out/r2aa_bp_plain.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -c 8000 --csv --log-file gpurun_out/r2aa_batch_launches.csv python profiles/batch_profile.py 64 32 4 c2 > gpurun_out/r2aa_bp_ncu.log 2>&1
python profiles/launch_summary.py gpurun_out/r2aa_batch_launches.csv 64 > gpurun_out/r2aa_batch_launch_summary.txt 2>&1
tail -20 gpurun_out/r2aa_batch_launch_summary.txt
