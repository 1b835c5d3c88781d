cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
run() { name=$1; shift; timeout -s KILL 120 python bench.py "$@" --no-cpu-baseline --no-e2e > gpurun_out/r18_$name.json 2> gpurun_out/r18_$name.err; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r18_$name.json")); r=d["roofline"]
    print("$name", "sigma", d["config"]["sigma"], "GFLOPS %.1f"%d["value"], "ms %.4f"%d["ms_per_step"], "kernel_ms %.4f"%r["kernel_ms_avg"], "frac %.3f"%r["frac"], "conv_ms %.2f"%d["config"]["csr_to_csr5_ms"], flush=True)
except Exception as e:
    print("$name FAILED", e); print(open("gpurun_out/r18_$name.err").read()[-400:])
PY
}
run c5_n1 --workload c5 --steps 50 --warmup 5
for s in 8 12 24 32; do run c2_s$s --steps 200 --warmup 10 --sigma $s; done
for s in 8 12 24 32; do run c3_s$s --workload c3 --steps 200 --warmup 10 --sigma $s; done
for s in 16 20 32; do run c4_s$s --workload c4 --steps 60 --warmup 5 --sigma $s; done
