"""Drop-in proof on the GPU: the REFERENCE's own benchmark driver
(CSR5_cuda/main.cu, compiled UNMODIFIED against include/anonymouslib_cuda.h by
tools/build_dropin_main.sh and linked with libcsr5_b200.so) runs a Matrix-Market file through
inputCSR / setX / setSigma / warmup / asCSR5 / spmv x (1 + 50 + NUM_RUN) / destroy and its own
end-of-main self-check (main.cu:360-384) must print `Check... PASS!`.

Because the reference calls spmv() 1051 times on the same y without re-zeroing and checks only the
first call, this also exercises repeated calls on the reference's own call sequence."""
import os
import re
import subprocess

import numpy as np
import pytest

from benchmark_spmv_using_csr5_b200 import matrices as M

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _write_mtx(path, A):
    """general real coordinate Matrix-Market file, 1-based (values are discarded by the reference:
    main.cu:314-326 overwrites them with rand() % 10)."""
    rows = np.repeat(np.arange(A.m), np.diff(A.row_ptr))
    with open(path, "w") as f:
        f.write("%%MatrixMarket matrix coordinate real general\n")
        f.write(f"{A.m} {A.n} {A.nnz}\n")
        np.savetxt(f, np.column_stack([rows + 1, A.col + 1, np.ones(A.nnz)]), fmt="%d %d %g")


@pytest.mark.parametrize("vt", ["double", "float"])
def test_reference_main_runs_against_our_library(tmp_path, vt):
    exe = os.path.join(ROOT, "oracle", "_ref", f"spmv_dropin_{vt}")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/spmv_dropin_* not built (tools/build_dropin_main.sh needs /root/reference)")
    A = M.example_c1()   # stand-in for the reference's missing example.mtx (README.md:27)
    mtx = str(tmp_path / "example.mtx")
    _write_mtx(mtx, A)
    r = subprocess.run([exe, mtx], capture_output=True, text=True, timeout=300)
    out = r.stdout
    assert r.returncode == 0, out + r.stderr
    assert "Check... PASS!" in out, out
    assert re.search(r"CSR5-based SpMV time = [\d.e+-]+ ms", out), out
    assert f"( {A.m}, {A.n} ) nnz = {A.nnz}" in out.replace("  ", " ") or str(A.nnz) in out, out
