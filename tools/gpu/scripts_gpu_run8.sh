cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
run() { name=$1; shift; timeout 300 python bench.py "$@" --no-cpu-baseline --no-e2e > gpurun_out/r8_$name.json 2> gpurun_out/r8_$name.err; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r8_$name.json")); r=d["roofline"]
    print("$name", "GFLOPS %.1f"%d["value"], "ms %.4f"%d["ms_per_step"], "kernel_ms %.4f"%r["kernel_ms_avg"], "frac %.3f"%r["frac"], r["kernel"], "hot", d["config"].get("hot_columns"), "%.3f"%d["config"].get("hot_coverage",0), flush=True)
except Exception as e:
    print("$name FAILED", e); print(open("gpurun_out/r8_$name.err").read()[-600:])
PY
}
for w in 2 4 8; do for n in 1 2 3 4; do
run c2_w${w}n${n} --steps 300 --warmup 20 --wpb $w --nch $n
done; done
for w in 2 4 8; do for n in 1 2 3 4; do
run c3_w${w}n${n} --workload c3 --hot 0 --steps 300 --warmup 20 --wpb $w --nch $n
done; done
run c3_w16n2 --workload c3 --hot 0 --steps 300 --warmup 20 --wpb 16 --nch 2
run c3_w16n4 --workload c3 --hot 0 --steps 300 --warmup 20 --wpb 16 --nch 4
for n in 2 3 4; do for t in 768 1024; do
run c3_hot16k_n${n}_t${t} --workload c3 --hot 16384 --steps 300 --warmup 20 --nch $n --hot-threads $t
done; done
run c3_hot12k_n3_t1024 --workload c3 --hot 12288 --steps 300 --warmup 20 --nch 3 --hot-threads 1024
for w in 2 4 8; do for n in 1 2 3 4; do
run c4_w${w}n${n} --workload c4 --steps 100 --warmup 10 --wpb $w --nch $n
done; done
timeout 300 python bench.py --steps 200 --warmup 20 > gpurun_out/r8_c2_full.json 2> gpurun_out/r8_c2_full.err; tail -3 gpurun_out/r8_c2_full.err; cat gpurun_out/r8_c2_full.json
