set -x
cd $GRAFT_REPO_ROOT
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/r2_multi_n${N}_smi.txt
if [ "$N" = "1" ]; then
  timeout 600 python benchmarks/c5_batch.py --pairs 4096 --lanes 8 --wave 32 --out gpurun_out/r02_c5_4096_n1.json > gpurun_out/r02_c5_4096_n1.log 2>&1
  timeout 600 python bench.py --gpus 1 --steps 100 --no-cpu-baseline --no-c1 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err
else
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 benchmarks/c5_batch.py --pairs 4096 --lanes 8 --wave 32 --out gpurun_out/r02_c5_4096_n${N}.json > gpurun_out/r02_c5_4096_n${N}.log 2>&1
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 100 --no-cpu-baseline --no-c1 > gpurun_out/r02_bench_n${N}.json 2> gpurun_out/r02_bench_n${N}.err
  timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k two_devices > gpurun_out/r02_two_devices_n${N}.log 2>&1
fi
tail -2 gpurun_out/r02_c5_4096_n${N}.log
