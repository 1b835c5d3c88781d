"""``python -m benchmark_spmv_using_csr5_b200.cli file.mtx`` -- the reference's benchmark driver
(``./spmv file.mtx``, CSR5_cuda/main.cu:119-396) on top of libcsr5_b200.so, printing the same report
lines: precision banner (main.cu:126-144), matrix line (:328), sequential CPU yardstick (:351-355),
device line (:30), ``CSR->CSR5 time`` (:76), ``CSR5-based SpMV time ... Bandwidth ... GFlops`` (:104-106)
and ``Check... PASS!`` (:360-384, 1 % relative tolerance against the scalar CSR loop).

Options mirror the reference's compile-time flags: ``--value-type double|float`` (VALUE_TYPE),
``--num-run N`` (NUM_RUN, default 1000); ``--results-csv`` appends ``filename,gflops`` like
CSR5_avx512/main.cpp:106-110; ``--seed`` makes the rand()%10 inputs reproducible.
"""
from __future__ import annotations

import argparse
import sys
import time

import numpy as np


def scalar_csr(m, row_ptr, col, val, x, alpha):
    """main.cu:336-350 (vectorised with numpy: per-row sums of x[col] * val * alpha)."""
    prod = x[col] * val * alpha
    cs = np.concatenate([[0], np.cumsum(prod, dtype=np.float64)])
    return (cs[row_ptr[1:]] - cs[row_ptr[:-1]]).astype(val.dtype)


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(prog="spmv")
    ap.add_argument("filename")
    ap.add_argument("--value-type", default="double", choices=["double", "float"])
    ap.add_argument("--num-run", type=int, default=1000)
    ap.add_argument("--sigma", type=int, default=-1)
    ap.add_argument("--seed", type=int, default=None)
    ap.add_argument("--results-csv", default=None)
    args = ap.parse_args(argv)

    import torch
    from . import handle as H
    from . import mmio

    np_dt, t_dt = (np.float64, torch.float64) if args.value_type == "double" else (np.float32, torch.float32)
    vb = np.dtype(np_dt).itemsize
    print("------------------------------------------------------")
    print(f"PRECISION = {'64-bit Double Precision' if vb == 8 else '32-bit Single Precision'}")
    print("------------------------------------------------------")
    print(f"--------------{args.filename}--------------")
    torch.cuda.set_device(0)
    try:
        # text parsing on the host, COO -> CSR (symmetric expansion + stable sort by row, main.cu:239-306) on the device
        m, n, d_rp, d_ci, _file_val = mmio.read_mtx_device(args.filename, np_dt)
    except mmio.MatrixMarketError as e:
        print(e)
        return 2
    row_ptr, col = d_rp.cpu().numpy(), d_ci.cpu().numpy()   # for the sequential CPU yardstick below
    nnz = len(col)
    val, x = mmio.reference_values(nnz, n, np_dt, args.seed)   # the file's values are discarded (main.cu:314-326)
    print(f" ( {m}, {n} ) nnz = {nnz}")
    gb = (m + 1 + nnz) * 4 + (2 * nnz + m) * vb                 # getB, detail/utils.h:10-14
    gflop = 2.0 * nnz
    alpha = 1.0

    t0 = time.perf_counter()
    y_ref = scalar_csr(m, row_ptr.astype(np.int64), col, val, x, np_dt(alpha))
    ref_ms = (time.perf_counter() - t0) * 1e3
    print(f"cpu sequential time = {ref_ms:.6g} ms. Bandwidth = {gb / (1e6 * ref_ms):.6g} GB/s. "
          f"GFlops = {gflop / (1e6 * ref_ms):.6g} GFlops.\n")

    prop = torch.cuda.get_device_properties(0)
    print(f"Device [0] {prop.name},  @ {getattr(prop, 'clock_rate', 0) * 1e-3:g}MHz. ")
    d_ci = d_ci.contiguous()
    d_val, d_x = torch.from_numpy(val).cuda(), torch.from_numpy(x).cuda()
    d_y = torch.zeros(m, device="cuda", dtype=t_dt)

    A = H.anonymouslibHandle(m, n, t_dt)
    A.inputCSR(nnz, d_rp, d_ci, d_val)
    A.setX(d_x)
    A.setSigma(args.sigma)
    A.warmup()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    err = A.asCSR5()
    e1.record()
    torch.cuda.synchronize()
    if err:
        print("asCSR5 err =", err, A.error_string(err))
        return 1
    print(f"omega = 32, sigma = {A.info().sigma}. ")
    print(f"CSR->CSR5 time = {e0.elapsed_time(e1):.6g} ms.")
    A.spmv(alpha, d_y)                                          # the checked call
    y = d_y.cpu().numpy()
    if args.num_run:
        for _ in range(50):
            A.spmv(alpha, d_y)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(args.num_run):
        A.spmv(alpha, d_y)
    e1.record()
    torch.cuda.synchronize()
    if args.num_run:
        t = e0.elapsed_time(e1) / args.num_run
        print(f"CSR5-based SpMV time = {t:.6g} ms. Bandwidth = {gb / (1e6 * t):.6g} GB/s. "
              f"GFlops = {gflop / (1e6 * t):.6g} GFlops.")
        if args.results_csv:
            with open(args.results_csv, "a") as f:
                f.write(f"{args.filename},{gflop / (1e6 * t):.6g}\n")
    A.destroy()
    A.free()

    errors = int(np.count_nonzero(np.abs(y_ref - y) > 0.01 * np.abs(y_ref)))
    print("Check... PASS!" if errors == 0 else f"Check... NO PASS! #Error = {errors} out of {m} entries.")
    print("------------------------------------------------------")
    return 0 if errors == 0 else 3


if __name__ == "__main__":
    sys.exit(main())
