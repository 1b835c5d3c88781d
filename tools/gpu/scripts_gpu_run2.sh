set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks.mem,power.draw --format=csv
nproc; lscpu | grep -E 'Model name|^CPU\(s\)'
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python bench.py --steps 200 --warmup 20 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; tail -2 gpurun_out/bench_c2.err; cat gpurun_out/bench_c2.json
python bench.py --steps 200 --warmup 20 --kernel 1 --no-cpu-baseline --no-e2e > gpurun_out/bench_c2_direct.json 2>> gpurun_out/bench_c2.err; cat gpurun_out/bench_c2_direct.json
python bench.py --workload c3 --steps 200 --warmup 20 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; tail -2 gpurun_out/bench_c3.err; cat gpurun_out/bench_c3.json
python bench.py --workload c3 --steps 200 --warmup 20 --kernel 1 --no-cpu-baseline --no-e2e > gpurun_out/bench_c3_direct.json 2>> gpurun_out/bench_c3.err; cat gpurun_out/bench_c3_direct.json
python bench.py --workload c4 --steps 100 --warmup 10 > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; tail -2 gpurun_out/bench_c4.err; cat gpurun_out/bench_c4.json
python bench.py --workload c4 --steps 100 --warmup 10 --kernel 1 --no-cpu-baseline --no-e2e > gpurun_out/bench_c4_direct.json 2>> gpurun_out/bench_c4.err; cat gpurun_out/bench_c4_direct.json
python bench.py --impl reference --steps 30 --warmup 5 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'csr5' -c 400 --csv --log-file gpurun_out/launches_c2.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'spmv_tma' -s 3 -c 2 -o gpurun_out/prof_c2_tma python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'spmv_direct' -s 3 -c 2 -o gpurun_out/prof_c2_direct python bench.py --steps 3 --warmup 3 --kernel 1 --no-cpu-baseline --no-e2e >> gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
