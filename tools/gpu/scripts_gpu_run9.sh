# 1 GPU: full GPU test suite, default benches for the three single-GPU configs, launch lists, ncu --set full
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 ) > gpurun_out/r9_pytest.log 2>&1; tail -4 gpurun_out/r9_pytest.log
( timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 )
timeout 400 python bench.py --steps 1000 --warmup 50 > gpurun_out/r9_bench_c2.json 2> gpurun_out/r9_bench_c2.err; tail -2 gpurun_out/r9_bench_c2.err; cat gpurun_out/r9_bench_c2.json
timeout 400 python bench.py --workload c3 --steps 1000 --warmup 50 > gpurun_out/r9_bench_c3.json 2> gpurun_out/r9_bench_c3.err; tail -2 gpurun_out/r9_bench_c3.err; cat gpurun_out/r9_bench_c3.json
timeout 400 python bench.py --workload c4 --steps 200 --warmup 20 > gpurun_out/r9_bench_c4.json 2> gpurun_out/r9_bench_c4.err; tail -2 gpurun_out/r9_bench_c4.err; cat gpurun_out/r9_bench_c4.json
timeout 300 python bench.py --impl reference --steps 30 --warmup 5 > gpurun_out/r9_bench_ref.json 2> gpurun_out/r9_bench_ref.err; cat gpurun_out/r9_bench_ref.json
K='regex:spmv_|calibrate|tile_|scan_|transpose|desc_offset|hot_|zero_rows'
for w in c2 c3 c4; do
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 400 --csv --log-file gpurun_out/r9_launches_$w.csv python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r9_ncu_launch_$w.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'spmv_direct' -s 3 -c 1 -o gpurun_out/r9_prof_${w}_direct python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r9_ncu_full_$w.log 2>&1
done
ls gpurun_out | grep r9_ | head -40
