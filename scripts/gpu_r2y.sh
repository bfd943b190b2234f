set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_batch.py tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r2y_test.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2y_test.log
tail -3 gpurun_out/r2y_test.log
for cfg in "8 64" "8 96" "8 128" "12 96" "6 64" "16 128"; do
  set -- $cfg
  timeout 300 python bench.py --steps 50 --no-cpu-baseline --no-c1 --no-extra-configs --batched-streams $1 --batched-wave $2 --batched-units 2048 > gpurun_out/r2y_b_$1_$2.json 2> gpurun_out/r2y_b_$1_$2.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2y_b_$1_$2.json').read().strip().splitlines()[-1])
    print('SWEEP lanes $1 wave $2', round(d['batched']['value']), round(d['batched']['e2e']['value']))
except Exception as e:
    print('SWEEP $1 $2 failed', e)
PY
done
