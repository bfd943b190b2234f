set -x
cd $GRAFT_REPO_ROOT
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/r2_multi_n${N}_smi.txt
if [ "$N" = "1" ]; then RUN="python"; else RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"; fi
for W in 64 128; do
  timeout 900 $RUN benchmarks/c5_batch.py --pairs 4096 --lanes 8 --wave $W --out gpurun_out/r02_c5_4096_n${N}_w${W}.json > gpurun_out/r02_c5_4096_n${N}_w${W}.log 2>&1
  tail -1 gpurun_out/r02_c5_4096_n${N}_w${W}.log
done
timeout 900 $RUN bench.py --gpus $N --steps 100 --no-cpu-baseline --no-c1 > gpurun_out/r02_bench_n${N}.json 2> gpurun_out/r02_bench_n${N}.err
if [ "$N" != "1" ]; then timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_batch.py -m gpu -q -k "two_devices or cpp" > gpurun_out/r02_two_devices_n${N}.log 2>&1; tail -1 gpurun_out/r02_two_devices_n${N}.log; fi
