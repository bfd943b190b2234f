set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for rep in 1 2 3; do
for v in "$@"; do
  export DDLO_GICP_LIB=$GRAFT_REPO_ROOT/dynamic_direct_lidar_odometry_b200/lib/variants/$v.so
  timeout 300 python bench.py --steps 20 --no-cpu-baseline --no-c1 --batched-streams 8 --batched-units 2048 > gpurun_out/ab_$v.json 2> gpurun_out/ab_$v.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/ab_$v.json').read().strip().splitlines()[-1])
    print('AB $v rep $rep', round(d['ms_per_step'],4), 'batched', round(d['batched'].get('value',0)), 'e2e-batched', round(d['batched'].get('e2e',{}).get('value',0)))
except Exception as e:
    print('AB $v failed', e)
PY
done
done
