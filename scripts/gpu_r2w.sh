set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2w_bench_n2.json 2> gpurun_out/r2w_bench_n2.err ) 2> gpurun_out/r2w_time.txt
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2w_ref_n2.json 2> gpurun_out/r2w_ref_n2.err ) 2>> gpurun_out/r2w_time.txt
timeout 600 python -m pytest tests -m gpu -x -q -k "two_devices or cpp_batch or multi" > gpurun_out/r2w_test.log 2>&1; tail -3 gpurun_out/r2w_test.log
