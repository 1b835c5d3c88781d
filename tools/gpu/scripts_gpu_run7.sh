cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 ) > gpurun_out/r7_pytest.log 2>&1
tail -25 gpurun_out/r7_pytest.log
run() { name=$1; shift; timeout 300 python bench.py "$@" --no-cpu-baseline --no-e2e > gpurun_out/r7_$name.json 2> gpurun_out/r7_$name.err; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r7_$name.json")); r=d["roofline"]
    print("$name", "GFLOPS %.1f"%d["value"], "ms %.4f"%d["ms_per_step"], "kernel_ms %.4f"%r["kernel_ms_avg"], "frac %.3f"%r["frac"], r["kernel"], "hot", d["config"].get("hot_columns"), "%.3f"%d["config"].get("hot_coverage",0), "conv_ms %.2f"%d["config"]["csr_to_csr5_ms"])
except Exception as e:
    print("$name FAILED", e); print(open("gpurun_out/r7_$name.err").read()[-800:])
PY
}
run c3_hot_auto --workload c3 --steps 200 --warmup 20
run c3_hot_off --workload c3 --steps 200 --warmup 20 --hot 0
run c3_hot_8k --workload c3 --steps 200 --warmup 20 --hot 8192
run c3_hot_24k --workload c3 --steps 200 --warmup 20 --hot 24576
run c3_hot_16k_t1024 --workload c3 --steps 200 --warmup 20 --hot-threads 1024
run c3_hot_16k_t512 --workload c3 --steps 200 --warmup 20 --hot-threads 512
run c4_default --workload c4 --steps 100 --warmup 10
run c4_w8n2 --workload c4 --steps 100 --warmup 10 --wpb 8 --nch 2
run c4_w4n1 --workload c4 --steps 100 --warmup 10 --wpb 4 --nch 1
run c4_w8n1 --workload c4 --steps 100 --warmup 10 --wpb 8 --nch 1
run c4_w4n3 --workload c4 --steps 100 --warmup 10 --wpb 4 --nch 3
run c4_w2n2 --workload c4 --steps 100 --warmup 10 --wpb 2 --nch 2
run c2_default --steps 300 --warmup 20
run c2_w8n1 --steps 300 --warmup 20 --wpb 8 --nch 1
run c2_w2n1 --steps 300 --warmup 20 --wpb 2 --nch 1
run c2_w4n2 --steps 300 --warmup 20 --wpb 4 --nch 2
run c3_w8n1 --workload c3 --steps 200 --warmup 20 --hot 0 --wpb 8 --nch 1
run c3_w4n2 --workload c3 --steps 200 --warmup 20 --hot 0 --wpb 4 --nch 2
