# compute-sanitizer over every kernel family on small cases (this GPU pool refuses the tool: profiles/r02_sanitizer_unavailable.txt),
# and the stand-in that did run: the -DDDLO_BOUNDS_CHECK build (device-side index checks, trap on the first violation).
#   gpurun --timeout 2400 -- 'bash scripts/gpu_sanitizer.sh'
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python profiles/sanitize_case.py > gpurun_out/sanitizer_$tool.txt 2>&1; echo "$tool rc=$?" >> gpurun_out/sanitizer_$tool.txt
done
( DDLO_NVCC_EXTRA=-DDDLO_BOUNDS_CHECK timeout 600 python -c "from dynamic_direct_lidar_odometry_b200 import build as b; b.build_library(force=True); print('bounds-checked build ok')" \
  && timeout 600 python profiles/sanitize_case.py && timeout 900 python -m pytest tests/test_gpu_batch.py tests/test_gpu_keyframes.py -m gpu -q \
  && timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "knn or covariances or linearize or align_small or voxel or stride or c2_align_matches_oracle" ) > gpurun_out/bounds_check.txt 2>&1
echo "bounds-check rc=$?" >> gpurun_out/bounds_check.txt
timeout 600 python -c "from dynamic_direct_lidar_odometry_b200 import build as b; b.build_library(force=True); print('regular build restored')" >> gpurun_out/bounds_check.txt 2>&1
tail -5 gpurun_out/bounds_check.txt
