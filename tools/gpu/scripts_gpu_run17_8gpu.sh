cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=8
run() { tag=$1; to=$2; shift 2
timeout -s KILL $to python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus $N "$@" > gpurun_out/r17_$tag.json 2> gpurun_out/r17_$tag.err
echo "bench $tag rc=$?"; grep -v "^\*\*\*\|OMP_NUM" gpurun_out/r17_$tag.err | tail -3
python - <<PY
import json
try:
    line=[l for l in open("gpurun_out/r17_$tag.json").read().splitlines() if l.startswith("{")][-1]
    d=json.loads(line); mg=d["multi_gpu"]; print("$tag", "GFLOPS %.1f"%d["value"], "ms %.4f"%d["ms_per_step"], "local %.4f nccl %.4f floor %.4f"%(mg["ms_per_step_spmv_only_no_exchange"], mg["ms_per_step_spmv_then_nccl_allgather"], mg["nvlink_time_floor_ms"]), mg["exchange"], "e2e", d["e2e"] and round(d["e2e"]["value"],1))
except Exception as e: print("FAILED", e)
PY
pkill -KILL -f bench.py; sleep 1
}
run c2_n8_auto 60 --steps 200 --warmup 20
run c5_n8_auto 85 --workload c5 --steps 50 --warmup 5 --no-e2e
