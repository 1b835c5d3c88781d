"""ctypes binding of libcsr5_b200.so (include/csr5_b200.h).

The product path has no CPU fallback: if the CUDA library is missing the import of any compute
entry point raises ``Csr5LibraryMissing`` -- it never routes through ``oracle/``.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
# CSR5B200_LIB: another build of the same library for A/B experiments (`make VARIANT=x EXTRA=-D...` -> libcsr5_b200_x.so)
LIB_PATH = os.path.join(_HERE, os.environ.get("CSR5B200_LIB", "libcsr5_b200.so"))
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "csr5_b200.h")
HEADER_PATHS = [HEADER_PATH, os.path.join(os.path.dirname(_HERE), "include", "csr5_b200_sharded.h")]


class Csr5LibraryMissing(RuntimeError):
    pass


class Csr5Info(C.Structure):
    """struct csr5b200_info of include/csr5_b200.h"""
    _fields_ = [
        ("format", C.c_int), ("m", C.c_int), ("n", C.c_int), ("nnz", C.c_int),
        ("value_bytes", C.c_int), ("sigma", C.c_int), ("bit_y_offset", C.c_int),
        ("bit_scansum_offset", C.c_int), ("num_packet", C.c_int), ("p", C.c_int),
        ("num_offsets", C.c_int), ("tail_partition_start", C.c_int), ("needs_zero_fill", C.c_int),
        ("kernel_in_use", C.c_int),
        ("partition_pointer", C.c_void_p), ("partition_descriptor", C.c_void_p),
        ("partition_descriptor_offset_pointer", C.c_void_p),
        ("partition_descriptor_offset", C.c_void_p), ("calibrator", C.c_void_p),
        ("last_cuda_error", C.c_int), ("launches_per_spmv", C.c_int),
        ("hot_columns", C.c_int), ("hot_coverage", C.c_double),
        ("convert_phase_ms", C.c_float * 8), ("convert_host_ms", C.c_double), ("convert_alloc_ms", C.c_double),
        ("exchange_transport", C.c_int), ("exchange_chunks", C.c_int), ("has_carries", C.c_int),
    ]


MAX_SCATTER = 8  # CSR5B200_MAX_SCATTER


class Csr5Exchange(C.Structure):
    """struct csr5b200_exchange of include/csr5_b200.h"""
    _fields_ = [
        ("rank", C.c_int), ("world", C.c_int),
        ("y_full", C.c_void_p * MAX_SCATTER), ("y_multicast", C.c_void_p),
        ("flags", C.c_void_p * MAX_SCATTER), ("row_begin", C.c_longlong),
        ("chunks", C.c_int), ("transport", C.c_int), ("entry_barrier", C.c_int),
        ("push_ctas", C.c_int), ("timeout_ms", C.c_int), ("push_threads", C.c_int),
    ]


# name -> (restype, argtypes); every symbol include/csr5_b200.h declares
SIGNATURES = {
    "csr5b200_create": (C.c_int, [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "csr5b200_warmup": (C.c_int, [C.c_void_p]),
    "csr5b200_input_csr": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "csr5b200_as_csr": (C.c_int, [C.c_void_p]),
    "csr5b200_as_csr5": (C.c_int, [C.c_void_p]),
    "csr5b200_set_x": (C.c_int, [C.c_void_p, C.c_void_p]),
    "csr5b200_spmv": (C.c_int, [C.c_void_p, C.c_double, C.c_void_p]),
    "csr5b200_spmv_axpby": (C.c_int, [C.c_void_p, C.c_double, C.c_double, C.c_void_p]),
    "csr5b200_spmv_allgather": (C.c_int, [C.c_void_p, C.c_double, C.c_double, C.POINTER(Csr5Exchange)]),
    "csr5b200_exchange_status": (C.c_int, [C.c_void_p]),
    "csr5b200_exchange_trace": (C.c_int, [C.c_void_p, C.POINTER(C.c_float), C.c_int, C.POINTER(C.c_int)]),
    "csr5b200_spmv_scatter": (C.c_int, [C.c_void_p, C.c_double, C.c_void_p, C.c_int, C.POINTER(C.c_void_p),
                                         C.c_int]),
    "csr5b200_destroy": (C.c_int, [C.c_void_p]),
    "csr5b200_set_sigma": (C.c_int, [C.c_void_p, C.c_int]),
    "csr5b200_free": (C.c_int, [C.c_void_p]),
    "csr5b200_set_stream": (C.c_int, [C.c_void_p, C.c_void_p]),
    "csr5b200_set_option": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "csr5b200_get_info": (C.c_int, [C.c_void_p, C.POINTER(Csr5Info)]),
    "csr5b200_get_kernel_times": (C.c_int, [C.c_void_p, C.POINTER(C.c_float), C.c_int, C.POINTER(C.c_int)]),
    "csr5b200_coo_to_csr": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                       C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_int), C.c_void_p]),
    "csr5b200_probe": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_float)]),
    "csr5b200_copy_meta_to_host": (C.c_int, [C.c_void_p] * 6),
    "csr5b200_spmv_host": (C.c_int, [C.c_void_p, C.c_double, C.c_void_p, C.c_void_p]),
    "csr5b200_spmv_host_batch": (C.c_int, [C.c_void_p, C.c_double, C.c_int, C.POINTER(C.c_void_p),
                                            C.POINTER(C.c_void_p)]),
    "csr5b200_call_anonymouslib": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                              C.c_void_p, C.c_void_p, C.c_void_p, C.c_double,
                                              C.c_int, C.c_int]),
    # include/csr5_b200_sharded.h
    "csr5b200_sharded_create": (C.c_int, [C.c_int, C.POINTER(C.c_int), C.c_int, C.POINTER(C.c_void_p)]),
    "csr5b200_sharded_input_csr_host": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                                   C.c_void_p]),
    "csr5b200_sharded_set_partition": (C.c_int, [C.c_void_p, C.c_double]),
    "csr5b200_sharded_set_sigma": (C.c_int, [C.c_void_p, C.c_int]),
    "csr5b200_sharded_set_option": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "csr5b200_sharded_set_exchange": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "csr5b200_sharded_set_x_host": (C.c_int, [C.c_void_p, C.c_void_p]),
    "csr5b200_sharded_as_csr5": (C.c_int, [C.c_void_p]),
    "csr5b200_sharded_spmv": (C.c_int, [C.c_void_p, C.c_double, C.c_double]),
    "csr5b200_sharded_iterate": (C.c_int, [C.c_void_p, C.c_int, C.c_double]),
    "csr5b200_sharded_synchronize": (C.c_int, [C.c_void_p]),
    "csr5b200_sharded_get_y": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_void_p)]),
    "csr5b200_sharded_copy_y_to_host": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "csr5b200_sharded_get_bounds": (C.c_int, [C.c_void_p, C.POINTER(C.c_longlong)]),
    "csr5b200_sharded_get_handle": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_void_p)]),
    "csr5b200_sharded_destroy": (C.c_int, [C.c_void_p]),
    "csr5b200_version": (C.c_char_p, []),
    "csr5b200_error_string": (C.c_char_p, [C.c_int]),
}

_lib = None


def build_library(force: bool = False) -> str:
    """Compile libcsr5_b200.so in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    csrc = os.path.join(_HERE, "csrc")
    if force:
        subprocess.run(["make", "-C", csrc, "clean"], check=True, capture_output=True)
    r = subprocess.run(["make", "-C", csrc, "-j4"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("building libcsr5_b200.so failed:\n" + r.stdout[-4000:] + r.stderr[-4000:])
    return LIB_PATH


def load_library():
    """dlopen the C-ABI library and type every declared entry point.  Fails loudly."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise Csr5LibraryMissing(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "or `make -C benchmark_spmv_using_csr5_b200/csrc` (there is no CPU fallback)")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the .so lacks a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib
