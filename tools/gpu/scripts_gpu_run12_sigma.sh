# 1 GPU: sigma sweep (is the Maxwell-era auto-sigma table still right on B200?)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
run() { name=$1; shift; timeout 300 python bench.py "$@" --no-cpu-baseline --no-e2e > gpurun_out/r12_$name.json 2> gpurun_out/r12_$name.err; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r12_$name.json")); r=d["roofline"]
    print("$name", "sigma", d["config"]["sigma"], "GFLOPS %.1f"%d["value"], "ms %.4f"%d["ms_per_step"], "kernel_ms %.4f"%r["kernel_ms_avg"], "frac %.3f"%r["frac"], "conv_ms %.2f"%d["config"]["csr_to_csr5_ms"], flush=True)
except Exception as e:
    print("$name FAILED", e); print(open("gpurun_out/r12_$name.err").read()[-400:])
PY
}
for s in 4 8 12 16 20 24 32; do run c2_s$s --steps 300 --warmup 20 --sigma $s; done
for s in 4 8 12 15 16 20 24 32; do run c3_s$s --workload c3 --steps 300 --warmup 20 --sigma $s; done
for s in 8 13 16 17 20 26 27 32; do run c4_s$s --workload c4 --steps 100 --warmup 10 --sigma $s; done
