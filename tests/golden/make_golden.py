"""Generates tests/golden/refcuda_*.npz: outputs of the REFERENCE's own CSR5_cuda backend
(oracle/_ref/libref_cuda.so, built by oracle/build_ref_cuda.sh from /root/reference) run on a B200.

Run on the GPU box (the reference CUDA code needs a GPU; /root/reference itself is not needed at run
time, only the prebuilt oracle/_ref/libref_cuda.so that travels with the snapshot):

    gpurun -- 'python tests/golden/make_golden.py gpurun_out/golden'

then copy gpurun_out/golden/*.npz into tests/golden/.  Inputs are NOT stored: they are regenerated from
the seeded generators of tests/cases.py (an input digest is stored and checked instead).  Each fixture
holds, for one (case, dtype): sigma/bit widths/p/num_offsets/tail_start, tile_ptr, tile_desc,
desc_offset_ptr, desc_offset, digests of the transposed col/val arrays, y after the FIRST spmv on a
zeroed y for integer-valued and for real-valued inputs, and y after THREE calls without re-zeroing
(documents the reference's accumulation drift).
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

# cases where the reference itself is undefined (OOB atomicOr) or has p = 1
# (zero-block launches) are left to the oracle-vs-scalar tests.
SKIP = {"trailing_empty_exact_multiple", "p1_tiny", "m1_short"}


def digest(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main(out_dir):
    import oracle
    from benchmark_spmv_using_csr5_b200 import matrices as M
    from tests.cases import small_cases
    os.makedirs(out_dir, exist_ok=True)
    for name, A, sigma in small_cases():
        if name in SKIP or A.nnz == 0:
            continue
        for dt, tag in ((np.float64, "f64"), (np.float32, "f32")):
            vi, xi = M.values(A.nnz, A.n, "int", dt)
            vr, xr = M.values(A.nnz, A.n, "real", dt)
            ri = oracle.ref_cuda_spmv(A.m, A.n, A.row_ptr, A.col, vi, xi, sigma, 1)
            rr = oracle.ref_cuda_spmv(A.m, A.n, A.row_ptr, A.col, vr, xr, sigma, 1)
            r3 = oracle.ref_cuda_spmv(A.m, A.n, A.row_ptr, A.col, vi, xi, sigma, 3)
            np.savez_compressed(
                os.path.join(out_dir, f"refcuda_{name}_{tag}.npz"),
                input_digest=digest(A.row_ptr) + digest(A.col) + digest(vi) + digest(xi),
                scalars=np.array([ri[k] for k in ("sigma", "bit_y", "bit_ss", "num_packet", "p",
                                                  "num_offsets", "tail_start")], np.int64),
                tile_ptr=ri["tile_ptr"], desc=ri["desc"], desc_off_ptr=ri["desc_off_ptr"],
                desc_off=ri["desc_off"], col5_digest=digest(ri["col5"]), val5_digest=digest(ri["val5"]),
                y_int=ri["y"], y_real=rr["y"], y_int_3calls=r3["y"])
            print("wrote", name, tag, "p =", ri["p"], "num_offsets =", ri["num_offsets"], flush=True)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "golden"))
