"""GPU tests against the REFERENCE's own CSR5_cuda backend run live on the same device
(oracle/_ref/libref_cuda.so, compat-patched build of /root/reference/CSR5_cuda made by
oracle/build_ref_cuda.sh; it travels to the GPU box as a prebuilt file).

north_star bar: "the reference's own CSR5_cuda y bit-exact on the same inputs" -- defined on the
reference's integer-valued input distribution (main.cu:314-326), where every summation order gives
the same bits; real-valued inputs are held to 1e-12 rel (FP64)."""
import numpy as np
import pytest

from benchmark_spmv_using_csr5_b200 import matrices as M
from tests.cases import small_cases
from tests.golden.make_golden import SKIP

pytestmark = pytest.mark.gpu
CASES = [c for c in small_cases() if c[0] not in SKIP]


@pytest.fixture(scope="module")
def refcuda(oracle):
    if not oracle.ref_cuda_available():
        pytest.skip("oracle/_ref/libref_cuda.so not built (needs /root/reference at build time)")
    import torch
    assert torch.cuda.is_available()
    return oracle


def _ours(A, val, x, sigma, kernel=0):
    import torch
    from benchmark_spmv_using_csr5_b200 import handle as H
    tdt = torch.float64 if val.dtype == np.float64 else torch.float32
    rp, ci = torch.from_numpy(A.row_ptr).cuda(), torch.from_numpy(A.col).cuda()
    v, xd = torch.from_numpy(val).cuda(), torch.from_numpy(x).cuda()
    h = H.anonymouslibHandle(A.m, A.n, tdt)
    assert h.inputCSR(A.nnz, rp, ci, v) == 0 and h.setX(xd) == 0
    h.setSigma(sigma)
    h.set_option(H.OPT_KERNEL, kernel if kernel < 3 else 0)
    h.set_option(H.OPT_HOT_COLUMNS, -1 if kernel == 3 else 0)   # table off: col stays the reference's
    assert h.asCSR5() == 0
    y = torch.zeros(A.m, device="cuda", dtype=tdt)
    assert h.spmv(1.0, y) == 0
    torch.cuda.synchronize()
    meta = h.meta_to_host()
    meta["col5"], meta["val5"] = ci.cpu().numpy(), v.cpu().numpy()
    h.free()
    return y.cpu().numpy(), meta


@pytest.mark.parametrize("dtype", [np.float64, np.float32], ids=["f64", "f32"])
@pytest.mark.parametrize("name,A,sigma", CASES, ids=[c[0] for c in CASES])
def test_against_reference_cuda(refcuda, name, A, sigma, dtype):
    val, x = M.values(A.nnz, A.n, "int", dtype)
    ref = refcuda.ref_cuda_spmv(A.m, A.n, A.row_ptr, A.col, val, x, sigma, 1)
    for kernel in (3, 1, 2):
        y, meta = _ours(A, val, x, sigma, kernel)
        assert np.array_equal(y, ref["y"]), f"{name}: y differs from the reference's CSR5_cuda (kernel {kernel})"
    for k in ("sigma", "bit_y", "bit_ss", "num_packet", "p", "tail_start"):
        assert meta[k] == ref[k], k
    p, npk = ref["p"], ref["num_packet"]
    assert np.array_equal(meta["tile_ptr"], ref["tile_ptr"]), "tile_ptr"
    live = (p - 1) * 32 * npk  # descriptors of tiles the reference's SpMV reads (t < p - 1)
    assert np.array_equal(meta["desc"][:live], ref["desc"][:live]), "tile_desc"
    any_dirty = bool((ref["tile_ptr"][:max(p - 1, 0)] >> 31).any())
    if any_dirty:
        # The reference's 4-tiles-per-block grid also counts the segments of the tail tile p - 1 (and
        # beyond) into num_offsets; nothing reads those entries.  What the SpMV reads
        # -- the exclusive scan up to tile p - 1 and the table entries below it -- must match exactly.
        assert np.array_equal(meta["desc_off_ptr"][:p], ref["desc_off_ptr"][:p]), "desc_offset_ptr"
        n = int(ref["desc_off_ptr"][p - 1])
        assert n <= meta["num_offsets"] <= ref["num_offsets"]
        assert np.array_equal(meta["desc_off"][:n], ref["desc_off"][:n]), "desc_offset"
    assert np.array_equal(meta["col5"], ref["col5"]), "transposed col"
    assert np.array_equal(meta["val5"], ref["val5"]), "transposed val"
    # real-valued inputs
    val, x = M.values(A.nnz, A.n, "real", dtype)
    ref = refcuda.ref_cuda_spmv(A.m, A.n, A.row_ptr, A.col, val, x, sigma, 1)
    y, _ = _ours(A, val, x, sigma, 3)
    rtol = 1e-12 if dtype == np.float64 else 2e-5
    assert np.allclose(y, ref["y"], rtol=rtol, atol=0 if dtype == np.float64 else 1e-5), name


def test_oracle_restatement_equals_reference_cuda(refcuda):
    """Pins oracle/csr5_oracle.c (the CPU restatement) to the code it restates."""
    for name, A, sigma in CASES:
        val, x = M.values(A.nnz, A.n, "int")
        ref = refcuda.ref_cuda_spmv(A.m, A.n, A.row_ptr, A.col, val, x, sigma, 1)
        assert np.array_equal(refcuda.csr5_spmv(A.m, A.n, A.row_ptr, A.col, val, x, sigma), ref["y"]), name
        want = refcuda.csr5_meta(A.m, A.nnz, ref["sigma"], A.row_ptr)
        assert np.array_equal(want.tile_ptr, ref["tile_ptr"]), name
        live = (ref["p"] - 1) * 32 * ref["num_packet"]
        assert np.array_equal(want.desc[:live], ref["desc"][:live]), name


def test_reference_drifts_on_repeated_calls_ours_does_not(refcuda):
    """the reference accumulates into rows at tile starts when y is not re-zeroed."""
    name, A, sigma = CASES[0]
    val, x = M.values(A.nnz, A.n, "int")
    once = refcuda.ref_cuda_spmv(A.m, A.n, A.row_ptr, A.col, val, x, sigma, 1)["y"]
    thrice = refcuda.ref_cuda_spmv(A.m, A.n, A.row_ptr, A.col, val, x, sigma, 3)["y"]
    assert not np.array_equal(once, thrice)
    y, _ = _ours(A, val, x, sigma)
    assert np.array_equal(y, once)
