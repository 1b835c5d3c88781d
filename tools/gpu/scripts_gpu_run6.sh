cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( timeout 400 python -m pytest tests/test_gpu_sharded.py -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/r6_pytest_sharded.log 2>&1
tail -5 gpurun_out/r6_pytest_sharded.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 200 --warmup 20 > gpurun_out/bench_c2_n2.json 2> gpurun_out/bench_c2_n2.err; tail -5 gpurun_out/bench_c2_n2.err; cat gpurun_out/bench_c2_n2.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 200 --warmup 20 --exchange nccl --no-e2e > gpurun_out/bench_c2_n2_nccl.json 2> gpurun_out/bench_c2_n2_nccl.err; tail -5 gpurun_out/bench_c2_n2_nccl.err; cat gpurun_out/bench_c2_n2_nccl.json
