# 8 GPUs: weak-scaling bench (C2 per GPU) with the fused exchange, strong-scaling C5 (R-MAT 25), sharded parity test
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=${1:-8}
nvidia-smi topo -m > gpurun_out/r11_topo_n$N.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29531 bench.py --gpus $N --steps 200 --warmup 20 > gpurun_out/r11_bench_c2_n$N.json 2> gpurun_out/r11_bench_c2_n$N.err; tail -3 gpurun_out/r11_bench_c2_n$N.err; cat gpurun_out/r11_bench_c2_n$N.json
timeout 400 $TR --master-port 29532 bench.py --gpus $N --workload c5 --steps 100 --warmup 10 > gpurun_out/r11_bench_c5_n$N.json 2> gpurun_out/r11_bench_c5_n$N.err; tail -3 gpurun_out/r11_bench_c5_n$N.err; cat gpurun_out/r11_bench_c5_n$N.json
( timeout 400 python -m pytest tests/test_gpu_sharded.py -m gpu -x -q 2>&1 | tail -5 ) > gpurun_out/r11_pytest_sharded_n$N.log 2>&1; tail -3 gpurun_out/r11_pytest_sharded_n$N.log
