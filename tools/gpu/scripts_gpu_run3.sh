set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python tests/golden/make_golden.py gpurun_out/golden 2>&1 | tail -5
python -m pytest tests/test_gpu_refcuda.py -m gpu -x -q 2>&1 | tail -15
python bench.py --workload c3 --steps 200 --warmup 20 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; tail -2 gpurun_out/bench_c3.err; cat gpurun_out/bench_c3.json
python bench.py --workload c3 --steps 200 --warmup 20 --kernel 2 --no-cpu-baseline --no-e2e > gpurun_out/bench_c3_tma.json 2>> gpurun_out/bench_c3.err; cat gpurun_out/bench_c3_tma.json
python bench.py --steps 1000 --warmup 50 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; tail -2 gpurun_out/bench_c2.err; cat gpurun_out/bench_c2.json
K='regex:spmv_|calibrate|tile_|scan_|transpose|desc_offset'
ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 400 --csv --log-file gpurun_out/launches_c2.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_launch.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 400 --csv --log-file gpurun_out/launches_c3.csv python bench.py --workload c3 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e >> gpurun_out/ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'spmv_direct' -s 3 -c 1 -o gpurun_out/prof_c3_direct python bench.py --workload c3 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'spmv_direct' -s 3 -c 1 -o gpurun_out/prof_c4_direct python bench.py --workload c4 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e >> gpurun_out/ncu_full.log 2>&1
python - <<'PY'
import sys, json, numpy as np
sys.path.insert(0, '.')
import oracle
from benchmark_spmv_using_csr5_b200 import matrices as M
out = {}
A = M.banded(10_000_000, 16)
val, x = M.values(A.nnz, A.n, "real", np.float64, 42)
ms, conv = oracle.ref_cuda_bench(A.m, A.n, A.row_ptr, A.col, val, x, -1, 50, 1000)
out["c2"] = {"ms": ms, "convert_ms": conv, "gflops": 2 * A.nnz / ms / 1e6}
print(out, flush=True)
A = M.rmat(20)
val, x = M.values(A.nnz, A.n, "real", np.float64, 42)
ms, conv = oracle.ref_cuda_bench(A.m, A.n, A.row_ptr, A.col, val, x, -1, 50, 1000)
out["rmat20"] = {"ms": ms, "convert_ms": conv, "gflops": 2 * A.nnz / ms / 1e6, "nnz": A.nnz}
print(out, flush=True)
json.dump(out, open("gpurun_out/refcuda_bench.json", "w"))
PY
ls gpurun_out gpurun_out/golden | head -80
