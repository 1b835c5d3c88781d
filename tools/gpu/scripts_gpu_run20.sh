cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout -s KILL 330 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/r20_pytest.log 2>&1; echo "pytest rc=$?"; tail -16 gpurun_out/r20_pytest.log | cut -c1-300
