cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=${1:-2}
timeout -s KILL 240 python -m pytest tests/test_gpu_sharded.py -m gpu -x -q > gpurun_out/r15_pytest_n$N.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r15_pytest_n$N.log | cut -c1-600
pkill -KILL -f multigpu_worker.py; sleep 1
run() { tag=$1; shift
timeout -s KILL 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus $N "$@" > gpurun_out/r15_$tag.json 2> gpurun_out/r15_$tag.err
echo "bench $tag rc=$?"; grep -v "^\*\*\*\|OMP_NUM" gpurun_out/r15_$tag.err | tail -4
python - <<PY
import json
try:
    line=[l for l in open("gpurun_out/r15_$tag.json").read().splitlines() if l.startswith("{")][-1]
    d=json.loads(line); mg=d["multi_gpu"]; print("$tag", "GFLOPS %.1f"%d["value"], "ms %.4f"%d["ms_per_step"], "local %.4f nccl %.4f floor %.4f"%(mg["ms_per_step_spmv_only_no_exchange"], mg["ms_per_step_spmv_then_nccl_allgather"], mg["nvlink_time_floor_ms"]), mg["exchange"], "e2e", d["e2e"] and round(d["e2e"]["value"],1))
except Exception as e: print("FAILED", e)
PY
pkill -KILL -f bench.py; sleep 1
}
if [ "$2" = "full" ]; then
run c2_n${N}_auto --steps 200 --warmup 20
run c2_n${N}_s2 --steps 100 --warmup 10 --no-e2e --scheme 2
run c2_n${N}_s1_uc --steps 100 --warmup 10 --no-e2e --scheme 1 --exchange fused-unicast
run c5_n${N}_auto --workload c5 --steps 50 --warmup 5
run c5_n${N}_s1 --workload c5 --steps 50 --warmup 5 --no-e2e --scheme 1
else
run c2_n${N}_s1 --steps 100 --warmup 10 --no-e2e --scheme 1
run c2_n${N}_s2 --steps 100 --warmup 10 --no-e2e --scheme 2
run c5_n${N}_s2 --workload c5 --steps 50 --warmup 5 --no-e2e --scheme 2
run c5_n${N}_s2_uc --workload c5 --steps 50 --warmup 5 --no-e2e --scheme 2 --exchange fused-unicast
fi
