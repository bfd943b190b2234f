set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for L in 4 8; do for W in 32 64 128; do
  timeout 300 python bench.py --steps 20 --no-cpu-baseline --no-c1 --batched-streams $L --batched-wave $W --batched-units 2048 > gpurun_out/r2o_l${L}_w$W.json 2>/dev/null
  python - <<PY
import json
d=json.loads(open('gpurun_out/r2o_l${L}_w$W.json').read().strip().splitlines()[-1])
print('COMBO lanes $L wave $W', round(d['batched']['value']), round(d['batched']['e2e']['value']))
PY
done; done
for W in 32 64 128; do timeout 300 python benchmarks/c5_batch.py --pairs 2048 --lanes 8 --wave $W > gpurun_out/r2o_c5_w$W.json 2>&1; python -c "
import json; d=json.loads(open('gpurun_out/r2o_c5_w$W.json').read().strip().splitlines()[-1]); print('C5 wave $W', round(d['value']))"; done
