set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for v in base "$@"; do
  if [ "$v" = "base" ]; then unset DDLO_GICP_LIB; else export DDLO_GICP_LIB=$GRAFT_REPO_ROOT/dynamic_direct_lidar_odometry_b200/lib/variants/$v.so; fi
  timeout 300 python bench.py --steps 100 --no-cpu-baseline --no-c1 --batched-streams 8 --batched-units 1024 > gpurun_out/var_$v.json 2> gpurun_out/var_$v.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/var_$v.json').read().strip().splitlines()[-1])
    print('VARIANT $v', round(d['ms_per_step'],4), {k: round(x,4) for k,x in d['stages_ms'].items()}, 'batched', round(d['batched'].get('value',0)), 'e2e-batched', round(d['batched'].get('e2e',{}).get('value',0)))
except Exception as e:
    print('VARIANT $v failed', e)
PY
done
