set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2i_test.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2i_test.log
tail -4 gpurun_out/r2i_test.log
timeout 600 python bench.py --steps 200 > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err
# the bounds-checked build (stands in for compute-sanitizer, which this pool refuses): every kernel family on small cases + the batch tests
( DDLO_NVCC_EXTRA=-DDDLO_BOUNDS_CHECK timeout 600 python -c "from dynamic_direct_lidar_odometry_b200 import build as b; b.build_library(force=True); print('bounds-checked build ok')" \
  && timeout 600 python profiles/sanitize_case.py && timeout 900 python -m pytest tests/test_gpu_batch.py tests/test_gpu_keyframes.py -m gpu -q \
  && timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "knn or covariances or linearize or align_small or voxel or stride or c2_align_matches_oracle" ) > gpurun_out/r2i_bounds_check.txt 2>&1
echo "bounds-check rc=$?" >> gpurun_out/r2i_bounds_check.txt
timeout 600 python -c "from dynamic_direct_lidar_odometry_b200 import build as b; b.build_library(force=True); print('regular build restored')" >> gpurun_out/r2i_bounds_check.txt 2>&1
tail -5 gpurun_out/r2i_bounds_check.txt
