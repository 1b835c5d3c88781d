# Experiments queued for the next round (not yet run; ~3 GPU-minutes on 1 GPU).  Usage under gpurun:
#   gpurun --timeout 400 -- 'bash tools/gpu/scripts_next_round.sh 2>&1 | tail -60'
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
run() { name=$1; shift; timeout -s KILL 90 python bench.py "$@" --no-cpu-baseline --no-e2e > gpurun_out/nx_$name.json 2> gpurun_out/nx_$name.err; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/nx_$name.json")); r=d["roofline"]
    print("$name", "GFLOPS %.1f"%d["value"], "ms %.4f"%d["ms_per_step"], "kernel_ms %.4f"%r["kernel_ms_avg"], "frac %.3f"%r["frac"], r["kernel"], flush=True)
except Exception as e:
    print("$name FAILED", e); print(open("gpurun_out/nx_$name.err").read()[-300:])
PY
}
# 1. TMA + x prefetch: ring geometry (warps per CTA x CTAs per SM x stages) on the three single-GPU configs
for w in 8 10 11 12; do run c2_k4_w$w --steps 200 --warmup 10 --kernel 4 --warps $w --ctas-per-sm 1; done
for st in 2 4; do run c2_k4_st$st --steps 200 --warmup 10 --kernel 4 --stages $st; done
for w in 8 10; do run c4_k4_w$w --workload c4 --steps 60 --warmup 5 --kernel 4 --warps $w --ctas-per-sm 1; done
run c4_k4_st2 --workload c4 --steps 60 --warmup 5 --kernel 4 --stages 2 --warps 15 --ctas-per-sm 1
# 2. C4 with the sigma the sweep preferred
run c4_s20 --workload c4 --steps 60 --warmup 5 --sigma 20
run c4_s20_k4 --workload c4 --steps 60 --warmup 5 --sigma 20 --kernel 4 --warps 12 --ctas-per-sm 1
# 3. hot-column table sizes with the 1024-thread CTA
for k in 4096 8192 12288; do run c3_hot$k --workload c3 --steps 200 --warmup 10 --hot $k --hot-threads 1024 --nch 3; done
