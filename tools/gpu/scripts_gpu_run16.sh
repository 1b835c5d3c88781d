# final 1-GPU validation of the tree as committed: full GPU test suite, smoke, default bench + reference arm
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests -m gpu -x -q > gpurun_out/r16_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r16_pytest.log | cut -c1-300
timeout -s KILL 100 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout -s KILL 200 python bench.py > gpurun_out/r16_bench_default.json 2> gpurun_out/r16_bench_default.err; echo "bench rc=$?"; tail -2 gpurun_out/r16_bench_default.err; cat gpurun_out/r16_bench_default.json | cut -c1-2600
