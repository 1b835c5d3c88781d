"""Floors of the SpMV kernel, measured (csr5b200_probe) next to the kernel itself, per workload; plus the
conversion phases per sigma and the hot-column variants.  Run on the GPU box: tools/gpu/run.sh probe[:c3,c5,c2].
Writes gpurun_out/<TAG>_probe.json; the table on stdout is what profiles/ keeps."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench  # noqa: E402  (workload builders, algorithmic bytes)
from benchmark_spmv_using_csr5_b200 import handle as H  # noqa: E402

TAG = os.environ.get("TAG", "r02")
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
peak, _ = bench.hbm_peak()
out = {}


def time_spmv(A, y, steps=60):
    for _ in range(5):
        A.spmv(1.0, y)
    A.kernel_times_ms()
    A.set_option(H.OPT_KERNEL_TIMING, 1)
    for _ in range(steps):
        A.spmv(1.0, y)
    kt = A.kernel_times_ms()
    A.set_option(H.OPT_KERNEL_TIMING, 0)
    return float(kt.mean()), float(kt.min())


for name in (sys.argv[1] if len(sys.argv) > 1 else "c3,c5,c2").split(","):
    w = bench.build_workload(name, torch, dev, 0, 1)
    m, n, nnz, dtype = w["m"], w["n"], w["col"].numel(), w["dtype"]
    vb = 8 if dtype == torch.float64 else 4
    b_alg = bench.algorithmic_bytes(m, n, nnz, vb)
    res = {"m": m, "n": n, "nnz": nnz, "b_alg": b_alg, "roofline_ms": b_alg / (peak * 1e6)}
    print(f"\n##### {name}: m={m} nnz={nnz} B_alg={b_alg/1e9:.3f} GB -> roofline {res['roofline_ms']:.4f} ms at {peak:.0f} GB/s")
    y = torch.empty(m, device=dev, dtype=dtype)

    def fresh(sigma=-1, hot=0, hot_threads=0, nch=0, wpb=0, kernel=0):
        A = H.anonymouslibHandle(m, n, dtype)
        assert A.inputCSR(nnz, w["row_ptr"], w["col"], w["val"]) == 0
        assert A.setX(w["x"]) == 0
        A.setSigma(sigma)
        A.set_option(H.OPT_HOT_COLUMNS, hot)
        A.set_option(H.OPT_HOT_THREADS, hot_threads)
        A.set_option(H.OPT_DIRECT_NCH, nch)
        A.set_option(H.OPT_DIRECT_WPB, wpb)
        A.set_option(H.OPT_KERNEL, kernel)
        return A

    # ---- conversion: phases per sigma, 1 cold + 4 warm conversions each (the 215 ms / 48 ms outliers of round 1) ----
    res["convert"] = {}
    sigmas = {"c3": [-1, 8, 12, 32], "c5": [-1], "c2": [-1, 8], "c4": [-1]}.get(name, [-1])
    for sg in sigmas:
        A = fresh(sg)
        rows = []
        for it in range(5):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            err = A.asCSR5()
            torch.cuda.synchronize()
            wall = (time.perf_counter() - t0) * 1e3
            assert err == 0
            i = A.info()
            rows.append({"wall_ms": wall, "host_ms": i.convert_host_ms, "alloc_ms": i.convert_alloc_ms,
                         "phases_ms": [round(float(v), 4) for v in list(i.convert_phase_ms)[:5]]})
            t0 = time.perf_counter()
            assert A.asCSR() == 0
            torch.cuda.synchronize()
            rows[-1]["as_csr_ms"] = (time.perf_counter() - t0) * 1e3
        sig = i.sigma
        res["convert"][str(sig)] = rows
        print(f"convert sigma={sig:2d} p={i.p}: " + " | ".join(
            f"{r['wall_ms']:.2f} ms (alloc {r['alloc_ms']:.2f}; tile_ptr/desc/scan/off/transpose {r['phases_ms']}; back {r['as_csr_ms']:.2f})"
            for r in (rows[0], rows[-1])))
        A.free()

    # ---- the kernel and its floors ------------------------------------------------------------------------------
    A = fresh()
    assert A.asCSR5() == 0
    k_avg, k_min = time_spmv(A, y)
    res["direct_ms"] = k_avg
    print(f"spmv_direct_kernel          {k_avg:.4f} ms (min {k_min:.4f})  frac {b_alg/(k_avg*1e6)/peak:.3f}")
    res["probe_ms"] = {}
    for pname, kind in H.PROBES.items():
        ms = A.probe(kind, 30)
        res["probe_ms"][pname] = ms
        print(f"  probe {pname:12s}        {ms:.4f} ms   (kernel / probe = {k_avg/ms:.3f})")
    gathers = (A.info().p - 1) * 32 * A.info().sigma
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    print(f"  gathers {gathers/1e6:.1f} M; at 1 divergent gather/clk/SM ({sms} SMs, 1.965 GHz): {gathers/sms/1.965e9*1e3:.4f} ms")
    A.free()

    # ---- variants ----------------------------------------------------------------------------------------------
    res["variants"] = {}
    if name in ("c3", "c5"):
        for label, kw in (("sigma14", dict(sigma=14)), ("sigma10", dict(sigma=10)),
                          ("nch1", dict(nch=1)), ("nch3", dict(nch=3)), ("wpb8", dict(wpb=8)), ("wpb2", dict(wpb=2)),
                          ("hot_auto", dict(hot=-1)), ("hot_4096_t1024", dict(hot=4096, hot_threads=1024)),
                          ("hot_8192_t1024", dict(hot=8192, hot_threads=1024)),
                          ("hot_16384_t768", dict(hot=16384, hot_threads=768)),
                          ("hot_24576_t1024", dict(hot=24576, hot_threads=1024))):
            try:
                A = fresh(**kw)
                t0 = time.perf_counter()
                assert A.asCSR5() == 0
                torch.cuda.synchronize()
                conv = (time.perf_counter() - t0) * 1e3
                k_avg, k_min = time_spmv(A, y, 40)
                i = A.info()
                res["variants"][label] = {"ms": k_avg, "convert_ms": conv, "hot_columns": i.hot_columns, "coverage": i.hot_coverage}
                print(f"variant {label:18s} {k_avg:.4f} ms  frac {b_alg/(k_avg*1e6)/peak:.3f}  (hot {i.hot_columns} cols, "
                      f"coverage {i.hot_coverage:.3f}, convert {conv:.1f} ms)")
                A.free()
            except Exception as e:
                print(f"variant {label}: {type(e).__name__}: {e}")
    out[name] = res
    del w, y
    torch.cuda.empty_cache()

os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"{TAG}_probe.json"), "w"), indent=1)
