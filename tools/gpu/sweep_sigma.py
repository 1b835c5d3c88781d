"""sigma sweep on B200: for nnz/row in 4..64 (banded, ~64 M nnz) and both value types, the SpMV time for every
candidate sigma next to the reference's choice (anonymouslib_cuda.h:297-313).  The table this prints is the basis of
CSR5B200_OPT_SIGMA_RULE = 1 (profiles/r02_sigma_rule.md).  Run: tools/gpu/run.sh sigma"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from benchmark_spmv_using_csr5_b200 import handle as H, matrices as M  # noqa: E402

TAG = os.environ.get("TAG", "r02")
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
CAND = [4, 5, 6, 7, 8, 10, 12, 14, 16, 18, 20, 22, 24, 26, 28, 30, 32]
NNZ = 64_000_000


def ref_sigma(k):
    return 4 if k <= 4 else (k if k <= 32 else (32 if k <= 256 else 6))


def kernel_ms(A, y, steps=30):
    for _ in range(4):
        A.spmv(1.0, y)
    A.kernel_times_ms()
    A.set_option(H.OPT_KERNEL_TIMING, 1)
    for _ in range(steps):
        A.spmv(1.0, y)
    kt = A.kernel_times_ms()
    A.set_option(H.OPT_KERNEL_TIMING, 0)
    return float(kt.mean())


def sweep(label, rp, ci, val, x, m, n, k, dtype, res):
    y = torch.empty(m, device=dev, dtype=dtype)
    row = {}
    for sg in sorted(set(CAND + [ref_sigma(k)])):
        A = H.anonymouslibHandle(m, n, dtype)
        assert A.inputCSR(ci.numel(), rp, ci, val) == 0
        assert A.setX(x) == 0
        A.setSigma(sg)
        err = A.asCSR5()
        if err:
            A.free()
            continue
        row[sg] = kernel_ms(A, y)
        A.free()
    best = min(row, key=row.get)
    ref = ref_sigma(k)
    res[label] = {"k": k, "ref_sigma": ref, "ms": row, "best_sigma": best, "gain_vs_ref": row[ref] / row[best]}
    top = sorted(row, key=row.get)[:4]
    print(f"{label:26s} k={k:3d} ref sigma {ref:2d}: {row[ref]*1e3:7.1f} us | best {best:2d}: {row[best]*1e3:7.1f} us "
          f"(x{row[ref]/row[best]:.3f}) | top4 {[(s, round(row[s]*1e3, 1)) for s in top]}", flush=True)


res = {}
for dtype, dn in ((torch.float64, "f64"), (torch.float32, "f32")):
    for k in (2, 3, 4, 5, 6, 8, 10, 12, 16, 20, 24, 27, 32, 40, 48, 64, 128, 300):
        m = NNZ // k
        rp, ci = M.device_banded(m, k, dev)
        val, x = M.device_values(ci.numel(), m, "real", dtype, dev, 42)
        sweep(f"banded_k{k}_{dn}", rp, ci, val, x, m, m, k, dtype, res)
        del rp, ci, val, x
        torch.cuda.empty_cache()
    rp, ci = M.device_rmat(21, device=dev)
    n = rp.numel() - 1
    val, x = M.device_values(ci.numel(), n, "real", dtype, dev, 42)
    sweep(f"rmat21_{dn}", rp, ci, val, x, n, n, ci.numel() // n, dtype, res)
    del rp, ci, val, x
    torch.cuda.empty_cache()
    rp, ci, val = M.device_laplacian27(160, device=dev, dtype=dtype)
    n = rp.numel() - 1
    _, x = M.device_values(1, n, "real", dtype, dev, 42)
    sweep(f"lap27_160_{dn}", rp, ci, val, x, n, n, ci.numel() // n, dtype, res)
    del rp, ci, val, x
    torch.cuda.empty_cache()
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", f"{TAG}_sweep_sigma.json"), "w"), indent=1)
