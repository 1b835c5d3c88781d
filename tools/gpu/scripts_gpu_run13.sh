cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
( timeout 400 python -m pytest tests/test_gpu_sharded.py -m gpu -x -q 2>&1 | tail -8 ) > gpurun_out/r13_pytest_sharded_n$N.log 2>&1; tail -6 gpurun_out/r13_pytest_sharded_n$N.log
p=29540
for ex in fused fused-unicast; do
p=$((p+1))
timeout 300 $TR --master-port $p bench.py --gpus $N --steps 200 --warmup 20 --exchange $ex --no-e2e > gpurun_out/r13_bench_c2_n${N}_$ex.json 2> gpurun_out/r13_bench_c2_n${N}_$ex.err; tail -3 gpurun_out/r13_bench_c2_n${N}_$ex.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r13_bench_c2_n${N}_$ex.json")); print("c2 N=$N $ex", "GFLOPS %.1f"%d["value"], "ms %.4f"%d["ms_per_step"], d["multi_gpu"])
except Exception as e: print("FAILED", e)
PY
done
if [ "$2" = "c5" ]; then
for ex in fused fused-unicast; do
p=$((p+1))
timeout 400 $TR --master-port $p bench.py --gpus $N --workload c5 --steps 100 --warmup 10 --exchange $ex --no-e2e > gpurun_out/r13_bench_c5_n${N}_$ex.json 2> gpurun_out/r13_bench_c5_n${N}_$ex.err; tail -3 gpurun_out/r13_bench_c5_n${N}_$ex.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r13_bench_c5_n${N}_$ex.json")); print("c5 N=$N $ex", "GFLOPS %.1f"%d["value"], "ms %.4f"%d["ms_per_step"], d["multi_gpu"])
except Exception as e: print("FAILED", e)
PY
done
fi
