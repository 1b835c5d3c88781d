"""CPU tests: pin the oracle (oracle/csr5_oracle.c, a restatement of the reference's CSR5_cuda)
against the reference's OWN CSR5_avx2 backend (oracle/_ref/libref_avx2.so, compiled from
/root/reference) and against the reference's scalar CSR yardstick (main.cu:336-350)."""
import numpy as np
import pytest

from benchmark_spmv_using_csr5_b200 import matrices as M
from tests.cases import sigma_sweep_case, small_cases

CASES = small_cases()


def _vals(A, kind, dtype=np.float64):
    return M.values(A.nnz, A.n, kind, dtype)


@pytest.mark.parametrize("name,A,sigma", CASES, ids=[c[0] for c in CASES])
def test_oracle_matches_scalar_int_bit_exact(oracle, name, A, sigma):
    val, x = _vals(A, "int")
    y = oracle.csr5_spmv(A.m, A.n, A.row_ptr, A.col, val, x, sigma)
    y_ref = oracle.csr_spmv(A.m, A.row_ptr, A.col, val, x)
    assert np.array_equal(y, y_ref)


@pytest.mark.parametrize("name,A,sigma", CASES, ids=[c[0] for c in CASES])
def test_oracle_matches_scalar_real(oracle, name, A, sigma):
    val, x = _vals(A, "real")
    y = oracle.csr5_spmv(A.m, A.n, A.row_ptr, A.col, val, x, sigma)
    y_ref = oracle.csr_spmv(A.m, A.row_ptr, A.col, val, x)
    assert np.allclose(y, y_ref, rtol=1e-12, atol=0)  # FP64: 1e-12 rel (north_star asks 1e-6)


@pytest.mark.parametrize("name,A,sigma", CASES, ids=[c[0] for c in CASES])
def test_oracle_matches_reference_avx2(oracle, name, A, sigma):
    """The reference's own AVX2 backend on the same inputs: bit-exact on its integer-valued
    input distribution, 1e-12 rel on real values."""
    if not oracle.ref_available():
        pytest.skip("oracle/_ref/libref_avx2.so not built")
    for kind in ("int", "real"):
        val, x = _vals(A, kind)
        y_ref = oracle.ref_avx2_spmv(A.m, A.n, A.row_ptr, A.col, val, x)
        y = oracle.csr5_spmv(A.m, A.n, A.row_ptr, A.col, val, x, sigma)
        if kind == "int":
            assert np.array_equal(y, y_ref)
        else:
            assert np.allclose(y, y_ref, rtol=1e-12, atol=0)


def test_oracle_fp32(oracle):
    for name, A, sigma in CASES[:8]:
        val, x = _vals(A, "int", np.float32)
        y = oracle.csr5_spmv(A.m, A.n, A.row_ptr, A.col, val, x, sigma)
        assert np.array_equal(y, oracle.csr_spmv(A.m, A.row_ptr, A.col, val, x)), name
        val, x = _vals(A, "real", np.float32)
        y = oracle.csr5_spmv(A.m, A.n, A.row_ptr, A.col, val, x, sigma)
        y64 = oracle.csr_spmv_f32_acc64(A.m, A.row_ptr, A.col, val, x)
        assert np.allclose(y, y64, rtol=2e-5, atol=1e-5), name  # FP32 tolerance


def test_oracle_all_sigmas(oracle):
    A = sigma_sweep_case()
    val, x = _vals(A, "int")
    y_ref = oracle.csr_spmv(A.m, A.row_ptr, A.col, val, x)
    for sigma in range(4, 33):
        y = oracle.csr5_spmv(A.m, A.n, A.row_ptr, A.col, val, x, sigma)
        assert np.array_equal(y, y_ref), sigma


def test_layout_scalars(oracle):
    # anonymouslib_cuda.h:121-137; SURVEY.md s8: C2 sigma 16 -> 1 packet (9+5+16), C4 sigma 26 -> 2 packets
    assert oracle.layout(16, 160_000_000) == (0, 9, 5, 1, 312_500)
    err, by, bs, npk, p = oracle.layout(26, 879_217_912)
    assert (err, by, bs, npk) == (0, 10, 5, 2) and p == -(-879_217_912 // (32 * 26))
    assert oracle.layout(4, 100)[1:4] == (7, 5, 1)
    assert oracle.layout(32, 10 ** 6)[1:4] == (10, 5, 2)


def test_auto_sigma_rule(oracle):
    # anonymouslib_cuda.h:297-313: r/s/t/u = 4/32/256/6
    assert oracle.auto_sigma(10, 30) == 4
    assert oracle.auto_sigma(10, 160) == 16
    assert oracle.auto_sigma(10, 320) == 32
    assert oracle.auto_sigma(10, 2560) == 32
    assert oracle.auto_sigma(10, 2570) == 6
    assert oracle.auto_sigma(1 << 22, 65_246_015) == 15


def test_transpose_round_trip(oracle):
    for name, A, sigma in CASES:
        s = sigma if sigma > 0 else oracle.auto_sigma(A.m, A.nnz)
        if A.nnz == 0:
            continue
        meta = oracle.csr5_meta(A.m, A.nnz, s, A.row_ptr)
        c5 = oracle.transpose(A.col, s, A.nnz, meta.tile_ptr, True)
        back = oracle.transpose(c5, s, A.nnz, meta.tile_ptr, False)
        assert np.array_equal(back, A.col), name


def test_tile_ptr_properties(oracle):
    for name, A, sigma in CASES:
        s = sigma if sigma > 0 else oracle.auto_sigma(A.m, A.nnz)
        meta = oracle.csr5_meta(A.m, A.nnz, s, A.row_ptr)
        rows = (meta.tile_ptr & oracle.MASK).astype(np.int64)
        assert np.all(np.diff(rows) >= 0), name
        assert rows[-1] == A.m, name
        for t in range(meta.p):
            b = min(t * 32 * s, A.nnz)
            r = rows[t]
            assert A.row_ptr[r] <= b and (r == A.m or A.row_ptr[r + 1] > b or r + 1 > A.m), name


# ---- randomised shapes (hypothesis): oracle == scalar CSR == the reference's CSR5_avx2 -------------
from hypothesis import HealthCheck, given, settings  # noqa: E402
from hypothesis import strategies as st  # noqa: E402


@st.composite
def _random_csr(draw):
    m = draw(st.integers(1, 400))
    n = draw(st.integers(1, 600))
    kind = draw(st.sampled_from(["short", "mixed", "hub", "sparse"]))
    seed = draw(st.integers(0, 2 ** 31 - 1))
    rng = np.random.default_rng(seed)
    if kind == "short":
        cnt = rng.integers(0, 6, size=m)
    elif kind == "mixed":
        cnt = rng.integers(0, 70, size=m)
    elif kind == "hub":
        cnt = rng.integers(0, 4, size=m)
        cnt[rng.integers(0, m)] = rng.integers(200, 4000)
    else:
        cnt = (rng.random(m) < 0.1) * rng.integers(1, 30, size=m)
    sigma = draw(st.sampled_from([-1, 4, 5, 7, 12, 16, 17, 26, 31, 32]))
    return M.from_row_counts(cnt, n, seed=seed + 1, name=f"{kind}_{m}x{n}"), sigma, seed


@settings(max_examples=60, deadline=None, suppress_health_check=[HealthCheck.too_slow, HealthCheck.function_scoped_fixture])
@given(case=_random_csr())
def test_oracle_random_shapes(oracle, case):
    A, sigma, seed = case
    if A.nnz == 0:
        return
    val, x = M.values(A.nnz, A.n, "int", np.float64, seed)
    y_ref = oracle.csr_spmv(A.m, A.row_ptr, A.col, val, x)
    assert np.array_equal(oracle.csr5_spmv(A.m, A.n, A.row_ptr, A.col, val, x, sigma), y_ref)
    if oracle.ref_available():
        assert np.array_equal(oracle.ref_avx2_spmv(A.m, A.n, A.row_ptr, A.col, val, x), y_ref)
    val, x = M.values(A.nnz, A.n, "real", np.float64, seed)
    y_ref = oracle.csr_spmv(A.m, A.row_ptr, A.col, val, x)
    assert np.allclose(oracle.csr5_spmv(A.m, A.n, A.row_ptr, A.col, val, x, sigma), y_ref, rtol=1e-12, atol=0)
