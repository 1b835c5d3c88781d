cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=${1:-2}
for mode in uc mc; do
timeout -s KILL 80 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) tools/gpu/debug_shard.py $mode > gpurun_out/debug_shard_$mode.out 2>&1
echo "mode $mode rc=$?"
grep -v "^\*\*\*\|OMP_NUM" gpurun_out/debug_shard_$mode.out | tail -5
grep -E "equal|done" gpurun_out/debug_shard_r0.log | cut -c1-120
cp gpurun_out/debug_shard_r0.log gpurun_out/debug_shard_${mode}_r0.log
pkill -KILL -f debug_shard.py; sleep 1
done
for ex in fused fused-unicast; do
timeout -s KILL 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus $N --steps 100 --warmup 10 --exchange $ex --no-e2e > gpurun_out/r14_bench_c2_n${N}_$ex.json 2> gpurun_out/r14_bench_c2_n${N}_$ex.err
echo "bench $ex rc=$?"; grep -v "^\*\*\*\|OMP_NUM" gpurun_out/r14_bench_c2_n${N}_$ex.err | tail -4
python - <<PY
import json
try:
    line=[l for l in open("gpurun_out/r14_bench_c2_n${N}_$ex.json").read().splitlines() if l.startswith("{")][-1]
    d=json.loads(line); print("c2 N=$N $ex", "GFLOPS %.1f"%d["value"], "ms %.4f"%d["ms_per_step"], d["multi_gpu"])
except Exception as e: print("FAILED", e)
PY
pkill -KILL -f bench.py; sleep 1
done
