cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout -s KILL 200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r22_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r22_pytest.log | cut -c1-400
run() { name=$1; shift; timeout -s KILL 60 python bench.py "$@" --no-cpu-baseline --no-e2e > gpurun_out/r22_$name.json 2> gpurun_out/r22_$name.err; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r22_$name.json")); r=d["roofline"]
    print("$name", "GFLOPS %.1f"%d["value"], "ms %.4f"%d["ms_per_step"], "kernel_ms %.4f"%r["kernel_ms_avg"], "frac %.3f"%r["frac"], r["kernel"], flush=True)
except Exception as e:
    print("$name FAILED", e); print(open("gpurun_out/r22_$name.err").read()[-300:])
PY
}
run c2_k4 --steps 200 --warmup 10 --kernel 4
run c4_k4 --workload c4 --steps 60 --warmup 5 --kernel 4
run c3_k4 --workload c3 --steps 200 --warmup 10 --kernel 4
run c2_k4_w6 --steps 200 --warmup 10 --kernel 4 --warps 11 --ctas-per-sm 1
