"""CPU oracle for the CSR5 SpMV hot path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package, and only as the checker / CPU baseline.
The product (``benchmark_spmv_using_csr5_b200``) never imports it.

Two libraries are wrapped with ctypes:

* ``libcsr5_oracle.so``  -- plain-C restatement of the reference's CSR5_cuda algorithm
  (``oracle/csr5_oracle.c``; every function cites the reference file:line it follows);
* ``_ref/libref_avx2.so`` -- the reference's OWN CSR5_avx2 backend compiled from
  ``/root/reference/CSR5_avx2`` by ``oracle/Makefile`` (no reference source is copied);
* ``_ref/libref_cuda.so`` -- the reference's OWN CSR5_cuda backend, compiled for sm_100a from a
  compat-patched scratch copy by ``oracle/build_ref_cuda.sh`` (needs a GPU to run).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
OMEGA = 32
MSB = 0x80000000
MASK = 0x7FFFFFFF

_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_u32p = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
_f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")


def build(force: bool = False) -> None:
    """Compile the checker (``make -C oracle``).  ``_ref`` is rebuilt only where the reference
    tree exists (the build container); elsewhere the prebuilt ``_ref/`` files are used."""
    so = os.path.join(_HERE, "libcsr5_oracle.so")
    src = os.path.join(_HERE, "csr5_oracle.c")
    ref = os.path.join(_HERE, "_ref", "libref_avx2.so")
    stale = force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src)
    need_ref = os.path.isdir("/root/reference/CSR5_avx2") and (
        force or not os.path.exists(ref)
        or os.path.getmtime(ref) < os.path.getmtime(os.path.join(_HERE, "ref_avx2_driver.cpp")))
    if stale or need_ref:
        if stale and os.path.exists(so):
            os.remove(so)
        subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)


_lib = None
_ref = None


def lib():
    """The C restatement (built on first use)."""
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(os.path.join(_HERE, "libcsr5_oracle.so"))
        L.csr5o_auto_sigma.argtypes = [C.c_int, C.c_int]
        L.csr5o_layout.argtypes = [C.c_int, C.c_int] + [C.POINTER(C.c_int)] * 4
        L.csr5o_tile_ptr.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, _i32p, _u32p]
        L.csr5o_tile_ptr.restype = None
        L.csr5o_tile_desc.argtypes = [C.c_int] * 6 + [_i32p, _u32p, _u32p, _i32p, C.POINTER(C.c_int)]
        L.csr5o_tile_desc.restype = None
        L.csr5o_desc_offset.argtypes = [C.c_int] * 5 + [_i32p, _u32p, _u32p, _i32p, _i32p]
        L.csr5o_desc_offset.restype = None
        L.csr5o_transpose.argtypes = [C.c_int, C.c_int, C.c_int, _u32p, C.c_void_p, C.c_int]
        L.csr5o_transpose.restype = None
        for name, vp in (("f64", _f64p), ("f32", _f32p)):
            f = getattr(L, f"csr5o_spmv_{name}")
            f.argtypes = [C.c_int] * 6 + [_i32p, _i32p, vp, _u32p, _u32p, _i32p, _i32p, vp, vp]
            f.restype = None
            g = getattr(L, f"csr5o_csr5_spmv_{name}")
            g.argtypes = [C.c_int] * 4 + [_i32p, _i32p, vp, vp, vp]
        L.csr5o_csr_spmv_f64.argtypes = [C.c_int, _i32p, _i32p, _f64p, _f64p, C.c_double, _f64p]
        L.csr5o_csr_spmv_f64.restype = None
        L.csr5o_csr_spmv_f32.argtypes = [C.c_int, _i32p, _i32p, _f32p, _f32p, C.c_float, _f32p]
        L.csr5o_csr_spmv_f32.restype = None
        L.csr5o_csr_axpby_f64.argtypes = [C.c_int, _i32p, _i32p, _f64p, _f64p, C.c_double, C.c_double, _f64p]
        L.csr5o_csr_axpby_f64.restype = None
        L.csr5o_csr_axpby_f32.argtypes = [C.c_int, _i32p, _i32p, _f32p, _f32p, C.c_float, C.c_float, _f32p]
        L.csr5o_csr_axpby_f32.restype = None
        L.csr5o_csr_spmv_f32_acc64.argtypes = [C.c_int, _i32p, _i32p, _f32p, _f32p, _f64p]
        L.csr5o_csr_spmv_f32_acc64.restype = None
        _lib = L
    return _lib


def ref_available() -> bool:
    return os.path.exists(os.path.join(_HERE, "_ref", "libref_avx2.so")) or \
        os.path.isdir("/root/reference/CSR5_avx2")


def ref():
    """The reference's own CSR5_avx2 backend (``oracle/_ref/libref_avx2.so``)."""
    global _ref
    if _ref is None:
        build()
        R = C.CDLL(os.path.join(_HERE, "_ref", "libref_avx2.so"))
        R.ref_avx2_spmv.argtypes = [C.c_int] * 3 + [_i32p, _i32p, _f64p, _f64p, _f64p, C.c_int]
        R.ref_avx2_bench.argtypes = [C.c_int] * 3 + [_i32p, _i32p, _f64p, _f64p, _f64p] + \
            [C.c_int] * 3 + [C.POINTER(C.c_double)] * 2
        R.ref_csr_omp_bench_f32.argtypes = [C.c_int, _i32p, _i32p, _f32p, _f32p, _f32p] + \
            [C.c_int] * 3 + [C.POINTER(C.c_double)]
        _ref = R
    return _ref


_refcuda = None


def ref_cuda_available() -> bool:
    return os.path.exists(os.path.join(_HERE, "_ref", "libref_cuda.so"))


def ref_cuda():
    """The reference's own CSR5_cuda backend (``oracle/_ref/libref_cuda.so``); GPU required."""
    global _refcuda
    if _refcuda is None:
        R = C.CDLL(os.path.join(_HERE, "_ref", "libref_cuda.so"))
        for name, vp in (("f64", _f64p), ("f32", _f32p)):
            f = getattr(R, f"ref_cuda_spmv_{name}")
            f.argtypes = [C.c_int] * 3 + [_i32p, _i32p, vp, vp, vp, C.c_int, C.c_int, _i32p,
                                          _u32p, C.c_size_t, _u32p, C.c_size_t, _i32p, C.c_size_t,
                                          _i32p, C.c_size_t, _i32p, vp]
            g = getattr(R, f"ref_cuda_bench_{name}")
            g.argtypes = [C.c_int] * 3 + [_i32p, _i32p, vp, vp] + [C.c_int] * 3 + [C.POINTER(C.c_double)] * 2
        _refcuda = R
    return _refcuda


def ref_cuda_spmv(m, n, row_ptr, col, val, x, sigma: int = -1, ncalls: int = 1) -> dict:
    """Runs the reference's CSR5_cuda (inputCSR/setX/setSigma/asCSR5/spmv on zeroed y) and returns y
    plus the CSR5 arrays its handle held.  Array capacities are sized from the CPU layout rules."""
    val = np.ascontiguousarray(val)
    name = {np.dtype(np.float64): "f64", np.dtype(np.float32): "f32"}[val.dtype]
    nnz = len(col)
    s = sigma if sigma > 0 else auto_sigma(m, nnz)
    err, by, bs, npk, p = layout(s, nnz)
    if err:
        raise ValueError(err)
    scal = np.zeros(8, np.int32)
    tile_ptr = np.zeros(p + 1, np.uint32)
    desc = np.zeros(max(p * OMEGA * npk, 1), np.uint32)
    dop = np.zeros(p + 1, np.int32)
    doff = np.zeros(m + p + 64, np.int32)
    col5 = np.zeros(max(nnz, 1), np.int32)
    val5 = np.zeros(max(nnz, 1), val.dtype)
    y = np.zeros(m, val.dtype)
    err = getattr(ref_cuda(), f"ref_cuda_spmv_{name}")(
        m, n, nnz, np.ascontiguousarray(row_ptr, np.int32), np.ascontiguousarray(col, np.int32), val,
        np.ascontiguousarray(x, val.dtype), y, sigma, ncalls, scal, tile_ptr, tile_ptr.size, desc, desc.size,
        dop, dop.size, doff, doff.size, col5, val5)
    if err:
        raise RuntimeError(f"reference CSR5_cuda returned {err}")
    return {"y": y, "sigma": int(scal[0]), "bit_y": int(scal[1]), "bit_ss": int(scal[2]),
            "num_packet": int(scal[3]), "p": int(scal[4]), "num_offsets": int(scal[5]),
            "tail_start": int(scal[6]), "tile_ptr": tile_ptr, "desc": desc[:p * OMEGA * npk],
            "desc_off_ptr": dop, "desc_off": doff[:int(scal[5])], "col5": col5[:nnz], "val5": val5[:nnz]}


def ref_cuda_bench(m, n, row_ptr, col, val, x, sigma: int = -1, warmup: int = 50, runs: int = 1000):
    """(ms per SpMV, conversion ms) of the reference's CSR5_cuda with the protocol of main.cu:79-106."""
    val = np.ascontiguousarray(val)
    name = {np.dtype(np.float64): "f64", np.dtype(np.float32): "f32"}[val.dtype]
    ms, conv = C.c_double(0), C.c_double(0)
    err = getattr(ref_cuda(), f"ref_cuda_bench_{name}")(
        m, n, len(col), np.ascontiguousarray(row_ptr, np.int32), np.ascontiguousarray(col, np.int32), val,
        np.ascontiguousarray(x, val.dtype), sigma, warmup, runs, C.byref(ms), C.byref(conv))
    if err:
        raise RuntimeError(f"reference CSR5_cuda returned {err}")
    return ms.value, conv.value


# ------------------------------------------------------------------------------------------------
# numpy-facing helpers
# ------------------------------------------------------------------------------------------------

@dataclass
class Csr5Meta:
    """CSR5 arrays exactly as the reference's CSR5_cuda handle would hold them
    (anonymouslib_cuda.h:25-52)."""
    sigma: int
    bit_y: int
    bit_ss: int
    num_packet: int
    p: int
    tile_ptr: np.ndarray       # (p+1,) uint32, bit 31 = tile has an empty row
    desc: np.ndarray           # (p*32*num_packet,) uint32
    desc_off_ptr: np.ndarray   # (p+1,) int32
    desc_off: np.ndarray       # (num_offsets,) int32
    num_offsets: int
    tail_start: int


def auto_sigma(m: int, nnz: int) -> int:
    return lib().csr5o_auto_sigma(m, nnz)


def layout(sigma: int, nnz: int):
    by, bs, npk, p = C.c_int(), C.c_int(), C.c_int(), C.c_int()
    err = lib().csr5o_layout(sigma, nnz, C.byref(by), C.byref(bs), C.byref(npk), C.byref(p))
    return err, by.value, bs.value, npk.value, p.value


def csr5_meta(m: int, nnz: int, sigma: int, row_ptr: np.ndarray) -> Csr5Meta:
    L = lib()
    row_ptr = np.ascontiguousarray(row_ptr, np.int32)
    err, by, bs, npk, p = layout(sigma, nnz)
    if err:
        raise ValueError(f"csr5o_layout error {err}")
    tile_ptr = np.zeros(p + 1, np.uint32)
    L.csr5o_tile_ptr(m, nnz, sigma, p, row_ptr, tile_ptr)
    desc = np.zeros(max(p * OMEGA * npk, 1), np.uint32)
    dop = np.zeros(p + 1, np.int32)
    nofs = C.c_int(0)
    L.csr5o_tile_desc(m, sigma, p, by, bs, npk, row_ptr, tile_ptr, desc, dop, C.byref(nofs))
    doff = np.zeros(max(nofs.value, 1), np.int32)
    if nofs.value:
        L.csr5o_desc_offset(sigma, p, by, bs, npk, row_ptr, tile_ptr, desc, dop, doff)
    return Csr5Meta(sigma, by, bs, npk, p, tile_ptr, desc[:p * OMEGA * npk], dop,
                    doff[:nofs.value], nofs.value, int(tile_ptr[p - 1] & MASK) if p else 0)


def transpose(arr: np.ndarray, sigma: int, nnz: int, tile_ptr: np.ndarray, r2c: bool) -> np.ndarray:
    out = np.ascontiguousarray(arr).copy()
    lib().csr5o_transpose(out.dtype.itemsize, sigma, nnz, np.ascontiguousarray(tile_ptr, np.uint32),
                          out.ctypes.data_as(C.c_void_p), 1 if r2c else 0)
    return out


def csr5_spmv(m, n, row_ptr, col, val, x, sigma: int = -1) -> np.ndarray:
    """y of the reference's CSR5_cuda algorithm (first call on zeroed y), CPU-emulated."""
    val = np.ascontiguousarray(val)
    name = {np.dtype(np.float64): "f64", np.dtype(np.float32): "f32"}[val.dtype]
    y = np.zeros(m, val.dtype)
    err = getattr(lib(), f"csr5o_csr5_spmv_{name}")(
        m, n, len(col), sigma, np.ascontiguousarray(row_ptr, np.int32),
        np.ascontiguousarray(col, np.int32), val, np.ascontiguousarray(x, val.dtype), y)
    if err:
        raise ValueError(f"oracle error {err}")
    return y


def csr_spmv(m, row_ptr, col, val, x, alpha: float = 1.0) -> np.ndarray:
    """The reference's scalar CSR yardstick (main.cu:336-350)."""
    val = np.ascontiguousarray(val)
    y = np.zeros(m, val.dtype)
    f = lib().csr5o_csr_spmv_f64 if val.dtype == np.float64 else lib().csr5o_csr_spmv_f32
    f(m, np.ascontiguousarray(row_ptr, np.int32), np.ascontiguousarray(col, np.int32), val,
      np.ascontiguousarray(x, val.dtype), alpha, y)
    return y


def csr_axpby(m, row_ptr, col, val, x, alpha: float, beta: float, y) -> np.ndarray:
    """alpha * A x + beta * y by the scalar CSR loop (the form stubbed at anonymouslib_cuda.h:281); returns a
    new array, y is not modified."""
    val = np.ascontiguousarray(val)
    out = np.ascontiguousarray(y, val.dtype).copy()
    f = lib().csr5o_csr_axpby_f64 if val.dtype == np.float64 else lib().csr5o_csr_axpby_f32
    f(m, np.ascontiguousarray(row_ptr, np.int32), np.ascontiguousarray(col, np.int32), val,
      np.ascontiguousarray(x, val.dtype), alpha, beta, out)
    return out


def csr_spmv_f32_acc64(m, row_ptr, col, val, x) -> np.ndarray:
    y = np.zeros(m, np.float64)
    lib().csr5o_csr_spmv_f32_acc64(m, np.ascontiguousarray(row_ptr, np.int32),
                                   np.ascontiguousarray(col, np.int32),
                                   np.ascontiguousarray(val, np.float32),
                                   np.ascontiguousarray(x, np.float32), y)
    return y


def ref_avx2_spmv(m, n, row_ptr, col, val, x, nthreads: int = 0) -> np.ndarray:
    """y of the reference's CSR5_avx2 backend (FP64, sigma 16, omega 4)."""
    y = np.zeros(m, np.float64)
    err = ref().ref_avx2_spmv(m, n, len(col), np.ascontiguousarray(row_ptr, np.int32),
                              np.ascontiguousarray(col, np.int32),
                              np.ascontiguousarray(val, np.float64),
                              np.ascontiguousarray(x, np.float64), y, nthreads)
    if err:
        raise RuntimeError(f"reference CSR5_avx2 returned {err}")
    return y


def ref_avx2_bench(m, n, row_ptr, col, val, x, nthreads=0, warmup=50, runs=100):
    """(ms per SpMV, conversion ms, y, threads) with the protocol of CSR5_avx2/main.cpp:41-79."""
    y = np.zeros(m, np.float64)
    ms, conv = C.c_double(0), C.c_double(0)
    err = ref().ref_avx2_bench(m, n, len(col), np.ascontiguousarray(row_ptr, np.int32),
                               np.ascontiguousarray(col, np.int32),
                               np.ascontiguousarray(val, np.float64),
                               np.ascontiguousarray(x, np.float64), y, nthreads, warmup, runs,
                               C.byref(ms), C.byref(conv))
    if err:
        raise RuntimeError(f"reference CSR5_avx2 returned {err}")
    threads = nthreads if nthreads > 0 else ref().ref_avx2_max_threads()
    return ms.value, conv.value, y, threads


def ref_csr_omp_bench_f32(m, row_ptr, col, val, x, nthreads=0, warmup=3, runs=10):
    y = np.zeros(m, np.float32)
    ms = C.c_double(0)
    ref().ref_csr_omp_bench_f32(m, np.ascontiguousarray(row_ptr, np.int32),
                                np.ascontiguousarray(col, np.int32),
                                np.ascontiguousarray(val, np.float32),
                                np.ascontiguousarray(x, np.float32), y, nthreads, warmup, runs,
                                C.byref(ms))
    threads = nthreads if nthreads > 0 else ref().ref_avx2_max_threads()
    return ms.value, y, threads
