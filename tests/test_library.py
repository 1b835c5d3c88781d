"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol
include/csr5_b200.h declares, and the host-side argument checks / error paths that need no GPU."""
import ctypes as C
import os
import re

import pytest

from benchmark_spmv_using_csr5_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = "".join(open(h).read() for h in _lib.HEADER_PATHS)   # csr5_b200.h + csr5_b200_sharded.h
    return sorted(set(re.findall(r"CSR5B200_API\s+[\w\s\*]+?\b(csr5b200_\w+)\s*\(", src)))


def test_library_builds_and_loads():
    _lib.build_library()
    lib = _lib.load_library()
    assert lib.csr5b200_version().decode().startswith("csr5-b200")


def test_every_declared_symbol_is_exported_and_typed():
    lib = _lib.load_library()
    declared = _declared_symbols()
    assert len(declared) >= 18
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/*.h but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes signature in _lib.SIGNATURES"
    assert sorted(_lib.SIGNATURES) == declared


def test_error_strings_and_codes():
    lib = _lib.load_library()
    # detail/common.h:13-18
    assert b"success" in lib.csr5b200_error_string(0)
    for code in (-1, -2, -3, -4, -5, -100, -101):
        assert lib.csr5b200_error_string(code)
    assert lib.csr5b200_create(4, 4, 2, C.byref(C.c_void_p())) == -5     # UNSUPPORTED_VALUE_TYPE
    assert lib.csr5b200_create(-1, 4, 8, C.byref(C.c_void_p())) == -101
    assert lib.csr5b200_spmv(None, 1.0, None) == -101


def test_handle_state_machine_without_gpu():
    """Call-order rules of anonymouslib_cuda.h that are decided on the host."""
    lib = _lib.load_library()
    h = C.c_void_p()
    assert lib.csr5b200_create(10, 10, 8, C.byref(h)) == 0
    assert lib.csr5b200_as_csr5(h) == -1                 # UNKOWN_FORMAT: inputCSR not called
    assert lib.csr5b200_spmv(h, 1.0, C.c_void_p(16)) == -1
    assert lib.csr5b200_input_csr(h, 30, None, None, None) == 0
    assert lib.csr5b200_spmv(h, 1.0, C.c_void_p(16)) == -4   # UNSUPPORTED_CSR_SPMV (anonymouslib_cuda.h:266-269)
    info = _lib.Csr5Info()
    assert lib.csr5b200_set_sigma(h, -1) == 0
    assert lib.csr5b200_get_info(h, C.byref(info)) == 0
    assert (info.m, info.n, info.nnz, info.sigma, info.format) == (10, 10, 30, 4, 0)   # 30 / 10 = 3 <= 4 -> 4
    assert lib.csr5b200_set_sigma(h, 3) == 0
    assert lib.csr5b200_as_csr5(h) == -3                 # CSR_TO_CSR5_FAILED: sigma outside [4, 32]
    assert lib.csr5b200_set_option(h, 99, 0) == -101
    assert lib.csr5b200_as_csr(h) == 0
    assert lib.csr5b200_free(h) == 0


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.Csr5LibraryMissing):
        _lib.load_library()


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "benchmark_spmv_using_csr5_b200")
    for dirpath, _dirs, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert not re.search(r"^\s*(import|from)\s+oracle\b", txt, re.M), f
                assert "csr5_oracle" not in txt and "libref_" not in txt, f


def test_header_is_plain_c_and_links(tmp_path):
    """include/csr5_b200.h is a C ABI: a C99 translation unit includes it, links against the library and
    calls entry points that need no GPU."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc") or "/usr/bin/gcc"
    src = tmp_path / "t.c"
    src.write_text('#include <stdio.h>\n#include "csr5_b200.h"\n#include "csr5_b200_sharded.h"\n'
                   'int main(void) { csr5b200_handle_t h = 0; csr5b200_info info; csr5b200_exchange ex; (void)ex;\n'
                   '  if (csr5b200_create(3, 4, 8, &h)) return 1;\n'
                   '  if (csr5b200_get_info(h, &info) || info.m != 3 || info.n != 4) return 2;\n'
                   '  if (csr5b200_spmv(h, 1.0, (void *)16) != CSR5B200_UNKNOWN_FORMAT) return 3;\n'
                   '  puts(csr5b200_version()); puts(csr5b200_error_string(CSR5B200_UNSUPPORTED_CSR_SPMV));\n'
                   '  return csr5b200_free(h); }\n')
    exe = tmp_path / "t"
    libdir = os.path.dirname(_lib.LIB_PATH)
    r = subprocess.run([gcc, "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                        "-L", libdir, "-lcsr5_b200", f"-Wl,-rpath,{libdir}"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0 and "csr5-b200" in r.stdout and "asCSR5" in r.stdout, r.stdout + r.stderr


def test_cpp_shim_compiles_standalone(tmp_path):
    """include/anonymouslib_cuda.h: the reference's call sequence (main.cu:59-108) compiles against the shim
    for both value types (compile only; the run is tests/test_gpu_dropin.py)."""
    import shutil
    import subprocess
    nvcc = shutil.which("nvcc")
    if not nvcc:
        pytest.skip("nvcc not found")
    src = tmp_path / "t.cu"
    src.write_text('#include "anonymouslib_cuda.h"\n'
                   'template <typename VT> int run(int m, int n, int nnz, int *rp, int *ci, VT *v, VT *x, VT *y) {\n'
                   '  anonymouslibHandle<int, unsigned int, VT> A(m, n);\n'
                   '  int err = A.inputCSR(nnz, rp, ci, v); err = A.setX(x);\n'
                   '  A.setSigma(ANONYMOUSLIB_AUTO_TUNED_SIGMA); A.warmup();\n'
                   '  anonymouslib_timer t; t.start(); err = A.asCSR5(); (void)t.stop();\n'
                   '  err = A.spmv((VT)1.0, y); A.destroy();\n'
                   '  double gb = getB<int, VT>(m, nnz), gf = getFLOP<int>(nnz); (void)gb; (void)gf;\n'
                   '  return err == ANONYMOUSLIB_SUCCESS ? 0 : err; }\n'
                   'template int run<double>(int, int, int, int *, int *, double *, double *, double *);\n'
                   'template int run<float>(int, int, int, int *, int *, float *, float *, float *);\n')
    r = subprocess.run([nvcc, "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-I",
                        os.path.join(ROOT, "include"), "-c", str(src), "-o", str(tmp_path / "t.o")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_round2_entry_points_reject_bad_arguments_without_a_gpu():
    """Host-side argument checks of the entry points added in round 2 (no compute call is made)."""
    lib = _lib.load_library()
    h = C.c_void_p()
    assert lib.csr5b200_create(3, 3, 8, C.byref(h)) == 0
    # y = alpha A x + beta y / the sharded step before inputCSR / asCSR5: the reference's format codes
    assert lib.csr5b200_spmv_axpby(h, 1.0, 1.0, C.c_void_p(16)) == -1            # ANONYMOUSLIB_UNKOWN_FORMAT
    ex = _lib.Csr5Exchange()
    ex.rank, ex.world = 0, 1
    assert lib.csr5b200_spmv_allgather(h, 1.0, 0.0, C.byref(ex)) == -1
    assert lib.csr5b200_spmv_allgather(None, 1.0, 0.0, C.byref(ex)) == -101
    assert lib.csr5b200_spmv_allgather(h, 1.0, 0.0, None) == -101
    assert lib.csr5b200_input_csr(h, 0, C.c_void_p(16), C.c_void_p(16), C.c_void_p(16)) == 0
    assert lib.csr5b200_spmv_axpby(h, 1.0, 1.0, C.c_void_p(16)) == -4            # spmv on CSR: asCSR5 first
    assert lib.csr5b200_spmv_allgather(h, 1.0, 0.0, C.byref(ex)) == -4
    for opt, bad in ((12, 2), (12, -1), (11, 3), (1, 3)):                          # sigma rule / exchange / kernel ids
        assert lib.csr5b200_set_option(h, opt, bad) == -101
    for opt in (12, 13, 14):                                                       # sigma rule, deterministic, trace
        assert lib.csr5b200_set_option(h, opt, 1) == 0
    ms, cnt = (C.c_float * 4)(), C.c_int(7)
    assert lib.csr5b200_exchange_trace(h, ms, 4, C.byref(cnt)) == 0 and cnt.value == 0
    assert lib.csr5b200_probe(h, 1, 0, ms) == -101 and lib.csr5b200_probe(h, 1, 1, ms) == -4
    assert lib.csr5b200_free(h) == 0
    # COO -> CSR
    n_out = C.c_int(0)
    assert lib.csr5b200_coo_to_csr(-1, 1, 0, None, None, None, 8, 0, C.c_void_p(16), None, None, 0, C.byref(n_out), None) == -101
    assert lib.csr5b200_coo_to_csr(1, 1, 1, None, None, None, 8, 0, C.c_void_p(16), None, None, 0, C.byref(n_out), None) == -101
    assert lib.csr5b200_coo_to_csr(1, 1, 0, None, None, None, 2, 0, C.c_void_p(16), None, None, 0, C.byref(n_out), None) == -5
    # sharded host API
    s = C.c_void_p()
    assert lib.csr5b200_sharded_create(0, (C.c_int * 1)(0), 8, C.byref(s)) == -101
    assert lib.csr5b200_sharded_create(9, (C.c_int * 9)(*([0] * 9)), 8, C.byref(s)) == -101
    assert lib.csr5b200_sharded_create(1, (C.c_int * 1)(0), 3, C.byref(s)) == -5
    assert lib.csr5b200_sharded_spmv(None, 1.0, 0.0) == -101
    assert lib.csr5b200_sharded_destroy(None) == 0
    assert b"timed out" in lib.csr5b200_error_string(-102)
