cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout -s KILL 200 python -m pytest tests/test_gpu_sharded.py -m gpu -x -q > gpurun_out/r19_pytest_n2.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r19_pytest_n2.log | cut -c1-600
pkill -KILL -f multigpu_worker.py; sleep 1
timeout -s KILL 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus 2 --steps 100 --warmup 10 > gpurun_out/r19_bench_c2_n2.json 2> gpurun_out/r19_bench_c2_n2.err; echo "bench n2 rc=$?"; grep -v "^\*\*\*\|OMP_NUM\|NCCL version" gpurun_out/r19_bench_c2_n2.err | tail -3; wc -l gpurun_out/r19_bench_c2_n2.json; cut -c1-1500 gpurun_out/r19_bench_c2_n2.json | head -3
pkill -KILL -f bench.py; sleep 1
CUDA_VISIBLE_DEVICES=0 timeout -s KILL 150 python bench.py --steps 100 --warmup 10 > gpurun_out/r19_bench_c2_n1.json 2> gpurun_out/r19_bench_c2_n1.err; echo "bench n1 rc=$?"; tail -2 gpurun_out/r19_bench_c2_n1.err; wc -l gpurun_out/r19_bench_c2_n1.json; python -c "
import json; d=json.load(open('gpurun_out/r19_bench_c2_n1.json')); print(d['value'], d['roofline']['frac'], d['e2e']['value'], d['config']['csr_to_csr5_ms'], d['config']['csr_to_csr5_ms_first_call'], d['cpu_baseline']['value'])"
