set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi topo -m | head -12
python -m pytest tests/test_gpu_sharded.py -m gpu -x -q 2>&1 | tail -15
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 200 --warmup 20 > gpurun_out/bench_c2_n2.json 2> gpurun_out/bench_c2_n2.err; tail -5 gpurun_out/bench_c2_n2.err; cat gpurun_out/bench_c2_n2.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 200 --warmup 20 --exchange nccl --no-e2e > gpurun_out/bench_c2_n2_nccl.json 2> gpurun_out/bench_c2_n2_nccl.err; tail -5 gpurun_out/bench_c2_n2_nccl.err; cat gpurun_out/bench_c2_n2_nccl.json
python bench.py --steps 200 --warmup 20 --no-cpu-baseline > gpurun_out/bench_c2_n1.json 2> gpurun_out/bench_c2_n1.err; tail -3 gpurun_out/bench_c2_n1.err; cat gpurun_out/bench_c2_n1.json
python bench.py --workload c3 --steps 200 --warmup 20 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; tail -3 gpurun_out/bench_c3.err; cat gpurun_out/bench_c3.json
