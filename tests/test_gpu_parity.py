"""GPU parity tests (-m gpu): the CUDA path, called through the C ABI (ctypes ->
libcsr5_b200.so), against the oracle on the same seeded inputs.

Bars: CSR5 metadata word-for-word; y bit-exact on the reference's integer-valued input
distribution (main.cu:314-326); FP64 real-valued y within 1e-12 rel (north_star: 1e-6);
FP32 real-valued within 2e-5 rel of an FP64-accumulated scalar loop."""
import numpy as np
import pytest

from benchmark_spmv_using_csr5_b200 import matrices as M
from tests.cases import sigma_sweep_case, small_cases

pytestmark = pytest.mark.gpu
CASES = small_cases()
FP64_RTOL = 1e-12
FP32_RTOL = 2e-5


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch


def _upload(torch, A, val, x):
    dev = "cuda"
    return (torch.from_numpy(A.row_ptr).to(dev), torch.from_numpy(A.col).to(dev),
            torch.from_numpy(val).to(dev), torch.from_numpy(x).to(dev))


# kernel ids: 1 direct-load, 2 TMA-staged (both with the hot-column table off), 3 hot-column kernel with
# the automatic table, 4 hot-column kernel with a tiny forced table (tagged and untagged columns mixed),
# 5 TMA-staged with the x gathers prefetched one tile ahead (CSR5B200_OPT_KERNEL = 4)
KERNELS = [1, 2, 3, 4, 5]
KERNEL_IDS = ["direct", "tma", "hot_auto", "hot_k48", "tma_prefetch"]
_OPT_KERNEL = {0: 0, 1: 1, 2: 2, 3: 0, 4: 0, 5: 4}
_OPT_HOT = {0: -1, 1: 0, 2: 0, 3: -1, 4: 48, 5: 0}


def _handle(torch, A, val, x, sigma, kernel=0):
    from benchmark_spmv_using_csr5_b200 import handle as H
    tdt = torch.float64 if val.dtype == np.float64 else torch.float32
    rp, ci, v, xd = _upload(torch, A, val, x)
    h = H.anonymouslibHandle(A.m, A.n, tdt)
    assert h.inputCSR(A.nnz, rp, ci, v) == 0
    assert h.setX(xd) == 0
    h.setSigma(sigma)
    assert h.set_option(H.OPT_KERNEL, _OPT_KERNEL[kernel]) == 0
    assert h.set_option(H.OPT_HOT_COLUMNS, _OPT_HOT[kernel]) == 0
    assert h.warmup() == 0
    assert h.asCSR5() == 0
    return h, (rp, ci, v, xd)


def _spmv(torch, h, m, dtype, alpha=1.0, fill=float("nan")):
    y = torch.full((m,), fill, device="cuda", dtype=dtype)  # garbage-filled: spmv must overwrite
    assert h.spmv(alpha, y) == 0
    torch.cuda.synchronize()
    return y.cpu().numpy()


@pytest.mark.parametrize("kernel", KERNELS, ids=KERNEL_IDS)
@pytest.mark.parametrize("name,A,sigma", CASES, ids=[c[0] for c in CASES])
def test_y_int_bit_exact_and_real(torch_cuda, oracle, name, A, sigma, kernel):
    torch = torch_cuda
    for kind in ("int", "real"):
        val, x = M.values(A.nnz, A.n, kind)
        h, _keep = _handle(torch, A, val, x, sigma, kernel)
        y = _spmv(torch, h, A.m, torch.float64)
        y_oracle = oracle.csr5_spmv(A.m, A.n, A.row_ptr, A.col, val, x, sigma)
        y_scalar = oracle.csr_spmv(A.m, A.row_ptr, A.col, val, x)
        if kind == "int":
            assert np.array_equal(y, y_oracle), f"{name}: differs from CSR5_cuda oracle"
            assert np.array_equal(y, y_scalar), f"{name}: differs from scalar CSR"
        else:
            assert np.allclose(y, y_oracle, rtol=FP64_RTOL, atol=0), name
            assert np.allclose(y, y_scalar, rtol=FP64_RTOL, atol=0), name
        assert h.destroy() == 0
        h.free()


@pytest.mark.parametrize("name,A,sigma", CASES, ids=[c[0] for c in CASES])
def test_metadata_word_for_word(torch_cuda, oracle, name, A, sigma):
    torch = torch_cuda
    if A.nnz == 0:
        pytest.skip("empty")
    val, x = M.values(A.nnz, A.n, "int")
    h, _keep = _handle(torch, A, val, x, sigma, kernel=1)   # hot-column table off: col is the reference's
    got = h.meta_to_host()
    s = sigma if sigma > 0 else oracle.auto_sigma(A.m, A.nnz)
    want = oracle.csr5_meta(A.m, A.nnz, s, A.row_ptr)
    assert (got["sigma"], got["bit_y"], got["bit_ss"], got["num_packet"], got["p"]) == \
        (want.sigma, want.bit_y, want.bit_ss, want.num_packet, want.p)
    assert got["tail_start"] == want.tail_start
    assert np.array_equal(got["tile_ptr"], want.tile_ptr), "tile_ptr"
    assert np.array_equal(got["desc"], want.desc), "tile_desc"
    assert got["num_offsets"] == want.num_offsets
    assert np.array_equal(got["desc_off_ptr"], want.desc_off_ptr), "desc_offset_ptr"
    assert np.array_equal(got["desc_off"], want.desc_off), "desc_offset"
    # transposed arrays as the reference's handle would hold them
    _rp, ci, v, _x = _keep
    torch.cuda.synchronize()
    assert np.array_equal(ci.cpu().numpy(), oracle.transpose(A.col, s, A.nnz, want.tile_ptr, True)), "col5"
    assert np.array_equal(v.cpu().numpy(), oracle.transpose(val, s, A.nnz, want.tile_ptr, True)), "val5"
    # asCSR restores the caller's arrays bit for bit
    assert h.asCSR() == 0
    torch.cuda.synchronize()
    assert np.array_equal(ci.cpu().numpy(), A.col)
    assert np.array_equal(v.cpu().numpy(), val)
    h.free()


@pytest.mark.parametrize("kernel", KERNELS, ids=KERNEL_IDS)
def test_all_sigmas(torch_cuda, oracle, kernel):
    torch = torch_cuda
    A = sigma_sweep_case()
    val, x = M.values(A.nnz, A.n, "int")
    y_ref = oracle.csr_spmv(A.m, A.row_ptr, A.col, val, x)
    for sigma in range(4, 33):
        h, _keep = _handle(torch, A, val, x, sigma, kernel)
        y = _spmv(torch, h, A.m, torch.float64)
        assert np.array_equal(y, y_ref), sigma
        h.free()


@pytest.mark.parametrize("kernel", KERNELS, ids=KERNEL_IDS)
def test_fp32(torch_cuda, oracle, kernel):
    torch = torch_cuda
    for name, A, sigma in CASES:
        val, x = M.values(A.nnz, A.n, "int", np.float32)
        h, _keep = _handle(torch, A, val, x, sigma, kernel)
        y = _spmv(torch, h, A.m, torch.float32)
        assert np.array_equal(y, oracle.csr_spmv(A.m, A.row_ptr, A.col, val, x)), name
        h.free()
        val, x = M.values(A.nnz, A.n, "real", np.float32)
        h, _keep = _handle(torch, A, val, x, sigma, kernel)
        y = _spmv(torch, h, A.m, torch.float32)
        y64 = oracle.csr_spmv_f32_acc64(A.m, A.row_ptr, A.col, val, x)
        assert np.allclose(y, y64, rtol=FP32_RTOL, atol=1e-5), name
        h.free()


def test_repeated_spmv_is_idempotent_and_alpha_honoured(torch_cuda, oracle):
    """The reference drifts on repeated calls and ignores alpha (s0-1); the
    replacement overwrites y and honours alpha, with a bug-compat switch for alpha."""
    torch = torch_cuda
    from benchmark_spmv_using_csr5_b200 import handle as H
    name, A, sigma = CASES[4]
    val, x = M.values(A.nnz, A.n, "int")
    h, _keep = _handle(torch, A, val, x, sigma)
    y = torch.zeros(A.m, device="cuda", dtype=torch.float64)
    for _ in range(5):
        assert h.spmv(1.0, y) == 0
    torch.cuda.synchronize()
    y_ref = oracle.csr_spmv(A.m, A.row_ptr, A.col, val, x)
    assert np.array_equal(y.cpu().numpy(), y_ref)
    assert np.array_equal(_spmv(torch, h, A.m, torch.float64, alpha=3.0), 3.0 * y_ref)
    h.set_option(H.OPT_IGNORE_ALPHA, 1)
    assert np.array_equal(_spmv(torch, h, A.m, torch.float64, alpha=3.0), y_ref)
    h.free()


def test_error_codes(torch_cuda):
    """detail/common.h:13-18 and the call-order rules of anonymouslib_cuda.h."""
    torch = torch_cuda
    from benchmark_spmv_using_csr5_b200 import handle as H
    A = M.banded(512, 16)
    val, x = M.values(A.nnz, A.n, "int")
    rp, ci, v, xd = _upload(torch, A, val, x)
    y = torch.zeros(A.m, device="cuda", dtype=torch.float64)
    h = H.anonymouslibHandle(A.m, A.n)
    assert h.asCSR5() == H.ANONYMOUSLIB_UNKOWN_FORMAT          # before inputCSR
    assert h.inputCSR(A.nnz, rp, ci, v) == 0
    h.setX(xd)
    assert h.spmv(1.0, y) == H.ANONYMOUSLIB_UNSUPPORTED_CSR_SPMV  # anonymouslib_cuda.h:266-269
    h.setSigma(3)
    assert h.asCSR5() == H.ANONYMOUSLIB_CSR_TO_CSR5_FAILED     # sigma outside [4, 32]
    h.setSigma(40)
    assert h.asCSR5() == H.ANONYMOUSLIB_CSR_TO_CSR5_FAILED
    h.setSigma(H.ANONYMOUSLIB_AUTO_TUNED_SIGMA)
    assert h.asCSR5() == 0 and h.asCSR5() == 0                 # second call is a no-op
    assert h.info().sigma == 16
    assert h.asCSR() == 0 and h.asCSR() == 0
    assert h.destroy() == 0
    h.free()
    with pytest.raises(TypeError):
        H.anonymouslibHandle(4, 4, torch.float16)


def test_host_entry_points(torch_cuda, oracle):
    torch = torch_cuda
    from benchmark_spmv_using_csr5_b200 import handle as H
    name, A, sigma = CASES[5]
    for dt in (np.float64, np.float32):
        val, x = M.values(A.nnz, A.n, "int", dt)
        y_ref = oracle.csr_spmv(A.m, A.row_ptr, A.col, val, x)
        y = H.call_anonymouslib(A.m, A.n, A.nnz, A.row_ptr, A.col, val, x, 1.0, sigma)
        assert np.array_equal(y, y_ref)
        h, _keep = _handle(torch, A, val, x, sigma)
        xp = torch.from_numpy(x).pin_memory()
        yp = torch.empty(A.m, dtype=xp.dtype).pin_memory()
        assert h.spmv_host(1.0, xp, yp) == 0
        assert np.array_equal(yp.numpy(), y_ref)
        y2 = np.empty(A.m, dt)
        assert h.spmv_host(2.0, x, y2) == 0
        assert np.array_equal(y2, 2 * y_ref)
        h.free()


def test_empty_matrix(torch_cuda):
    torch = torch_cuda
    from benchmark_spmv_using_csr5_b200 import handle as H
    m = 100
    rp = torch.zeros(m + 1, device="cuda", dtype=torch.int32)
    ci = torch.zeros(1, device="cuda", dtype=torch.int32)
    v = torch.zeros(1, device="cuda", dtype=torch.float64)
    x = torch.ones(m, device="cuda", dtype=torch.float64)
    h = H.anonymouslibHandle(m, m)
    assert h.inputCSR(0, rp, ci, v) == 0
    h.setX(x)
    h.setSigma(-1)
    assert h.asCSR5() == 0
    y = torch.full((m,), 7.0, device="cuda", dtype=torch.float64)
    assert h.spmv(1.0, y) == 0
    torch.cuda.synchronize()
    assert float(y.abs().sum()) == 0.0
    h.free()


def test_non_default_stream(torch_cuda, oracle):
    torch = torch_cuda
    name, A, sigma = CASES[0]
    val, x = M.values(A.nnz, A.n, "int")
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        h, _keep = _handle(torch, A, val, x, sigma)
        y = torch.empty(A.m, device="cuda", dtype=torch.float64)
        assert h.spmv(1.0, y) == 0
    s.synchronize()
    assert np.array_equal(y.cpu().numpy(), oracle.csr_spmv(A.m, A.row_ptr, A.col, val, x))
    h.free()


# ---- BASELINE.json full sizes: size-independent properties, evaluated on the device ------------

def _segment_sums_exact(torch, rp, ci, v, x):
    """Exact y for integer-valued inputs: prefix sums of the products are exact in FP64 below 2^53."""
    prod = v.double() * x.double()[ci.long()]
    cs = torch.zeros(prod.numel() + 1, device=prod.device, dtype=torch.float64)
    torch.cumsum(prod, 0, out=cs[1:])
    return cs[rp[1:].long()] - cs[rp[:-1].long()]


def _full_size_check(torch, rp, ci, v, x, dtype, kernels=(1, 2, 3, 5)):
    from benchmark_spmv_using_csr5_b200 import handle as H
    m, n, nnz = rp.numel() - 1, x.numel(), ci.numel()
    y_ref = _segment_sums_exact(torch, rp, ci, v, x).to(dtype)
    col0, val0 = ci.clone(), v.clone()
    for kernel in kernels:
        h = H.anonymouslibHandle(m, n, dtype)
        assert h.inputCSR(nnz, rp, ci, v) == 0
        h.setX(x)
        h.setSigma(-1)
        h.set_option(H.OPT_KERNEL, _OPT_KERNEL[kernel])
        h.set_option(H.OPT_HOT_COLUMNS, _OPT_HOT[kernel])
        assert h.asCSR5() == 0
        y = torch.full((m,), float("nan"), device="cuda", dtype=dtype)
        assert h.spmv(1.0, y) == 0
        assert torch.equal(y, y_ref), f"kernel {kernel}: y differs from exact segment sums"
        # linearity: A(2x) == 2 A x exactly for integer data
        x2 = (2 * x).contiguous()
        h.setX(x2)
        y2 = torch.empty_like(y)
        assert h.spmv(1.0, y2) == 0
        assert torch.equal(y2, 2 * y_ref)
        h.setX(x)
        assert h.destroy() == 0   # round trip restores the caller's arrays
        assert torch.equal(ci, col0) and torch.equal(v, val0)
        h.free()


def test_full_size_c2_banded(torch_cuda):
    """BASELINE.json configs[1]: banded 10M x 10M, 16 nnz/row, FP64."""
    torch = torch_cuda
    m = 10_000_000
    rp, ci = M.device_banded(m, 16)
    v, x = M.device_values(ci.numel(), m, "int", torch.float64, "cuda")
    _full_size_check(torch, rp, ci, v, x, torch.float64)


def test_full_size_c3_rmat22(torch_cuda):
    """BASELINE.json configs[2]: R-MAT scale 22, ~64M nnz, FP64 (52 % empty rows, hub rows)."""
    torch = torch_cuda
    rp, ci = M.device_rmat(22)
    n = rp.numel() - 1
    v, x = M.device_values(ci.numel(), n, "int", torch.float64, "cuda")
    _full_size_check(torch, rp, ci, v, x, torch.float64)


def test_hot_column_table(torch_cuda, oracle):
    """The hot-column table (no reference counterpart): chosen at asCSR5(), tags col in place, asCSR()
    restores col bit for bit; auto mode drops it when it would serve < 25 % of the references."""
    torch = torch_cuda
    from benchmark_spmv_using_csr5_b200 import handle as H
    A = M.rmat(14)
    val, x = M.values(A.nnz, A.n, "int")
    y_ref = oracle.csr_spmv(A.m, A.row_ptr, A.col, val, x)
    for cap in (-1, 16, 1000, 100000):
        h, keep = _handle(torch, A, val, x, -1, kernel=1)
        h.asCSR()
        h.set_option(H.OPT_HOT_COLUMNS, cap)
        assert h.asCSR5() == 0
        info = h.info()
        assert info.hot_columns > 0 and 0.0 < info.hot_coverage <= 1.0
        if cap > 0:
            assert info.hot_columns <= cap
        tagged = (keep[1].cpu().numpy() < 0)
        live = (info.p - 1) * 32 * info.sigma
        assert tagged[:live].any() and not tagged[live:].any()          # the tail tile is never tagged
        assert abs(tagged[:live].mean() - info.hot_coverage) < 1e-9
        y = _spmv(torch, h, A.m, torch.float64)
        assert h.info().kernel_in_use == H.KERNEL_HOT
        assert np.array_equal(y, y_ref), cap
        assert h.asCSR() == 0
        torch.cuda.synchronize()
        assert np.array_equal(keep[1].cpu().numpy(), A.col) and np.array_equal(keep[2].cpu().numpy(), val)
        h.free()
    # banded: every column is referenced equally often and there are more of them than slots -> no table
    B = M.banded(200_000, 16)
    val, x = M.values(B.nnz, B.n, "int")
    h, keep = _handle(torch, B, val, x, -1, kernel=0)
    assert h.info().hot_columns == 0
    y = _spmv(torch, h, B.m, torch.float64)
    assert h.info().kernel_in_use == H.KERNEL_DIRECT
    assert np.array_equal(y, oracle.csr_spmv(B.m, B.row_ptr, B.col, val, x))
    h.free()


def test_full_size_c4_laplacian_fp32(torch_cuda):
    """BASELINE.json configs[3]: 27-pt Laplacian on 320^3 (879 M nnz, FP32, sigma 26 -> 2-packet descriptors,
    byte offsets beyond 4 GiB).  Values are 26 / -1 and x is integer-valued, so every row sum is an integer
    below 2^24: FP32 results must equal the FP64 prefix-sum segment sums bit for bit."""
    torch = torch_cuda
    rp, ci, v = M.device_laplacian27(320, dtype=torch.float32)
    n = rp.numel() - 1
    _, x = M.device_values(1, n, "int", torch.float32, "cuda")
    _full_size_check(torch, rp, ci, v, x, torch.float32, kernels=(1, 2, 5))


def test_spmv_host_batch_pipeline(torch_cuda, oracle):
    """csr5b200_spmv_host_batch: K independent SpMVs on host vectors, pipelined; every y_k must equal the
    oracle's, including when buffers rotate (the double-buffered staging is reused every second vector)."""
    torch = torch_cuda
    name, A, sigma = CASES[4]
    for dt, tdt in ((np.float64, torch.float64), (np.float32, torch.float32)):
        val, _ = M.values(A.nnz, A.n, "int", dt)
        h, _keep = _handle(torch, A, val, np.zeros(A.n, dt), sigma)
        xs, ys, want = [], [], []
        for k in range(7):
            x = np.random.default_rng(k).integers(0, 10, A.n).astype(dt)
            xs.append(torch.from_numpy(x).pin_memory())
            ys.append(torch.full((A.m,), float("nan"), dtype=tdt).pin_memory())
            want.append(oracle.csr_spmv(A.m, A.row_ptr, A.col, val, x))
        assert h.spmv_host_batch(1.0, xs, ys) == 0
        for k in range(7):
            assert np.array_equal(ys[k].numpy(), want[k]), (dt, k)
        assert h.spmv_host_batch(2.0, xs[:1], ys[:1]) == 0
        assert np.array_equal(ys[0].numpy(), 2 * want[0])
        assert h.spmv_host_batch(1.0, [], []) == 0
        h.free()


def test_random_shapes_all_kernels(torch_cuda, oracle):
    """40 seeded random row-length profiles x random sigma through every kernel variant."""
    torch = torch_cuda
    rng = np.random.default_rng(2024)
    for trial in range(40):
        m, n = int(rng.integers(1, 3000)), int(rng.integers(1, 5000))
        kind = trial % 4
        if kind == 0:
            cnt = rng.integers(0, 6, size=m)
        elif kind == 1:
            cnt = rng.integers(0, 70, size=m)
        elif kind == 2:
            cnt = rng.integers(0, 4, size=m)
            cnt[rng.integers(0, m)] = rng.integers(200, 30000)
        else:
            cnt = (rng.random(m) < 0.1) * rng.integers(1, 30, size=m)
        A = M.from_row_counts(cnt, n, seed=trial)
        if A.nnz == 0:
            continue
        sigma = int(rng.choice([-1, 4, 5, 7, 12, 16, 17, 26, 31, 32]))
        dt, tdt = ((np.float64, torch.float64), (np.float32, torch.float32))[trial % 2]
        val, x = M.values(A.nnz, A.n, "int", dt, seed=trial)
        y_ref = oracle.csr_spmv(A.m, A.row_ptr, A.col, val, x)
        for kernel in KERNELS:
            h, _keep = _handle(torch, A, val, x, sigma, kernel)
            y = _spmv(torch, h, A.m, tdt)
            assert np.array_equal(y, y_ref), (trial, kernel, sigma, m, n)
            h.free()


# ---- round 2: real-valued full-size parity, the reference's AVX2 path directly, C5 on one GPU -----------------

def _row_sums_fp64(torch, rp, ci, v, x):
    """Per-row FP64 sums and their magnitudes by an independent route (index_add_ of the products onto their row)."""
    m = rp.numel() - 1
    counts = (rp[1:] - rp[:-1]).long()
    rows = torch.repeat_interleave(torch.arange(m, device=rp.device), counts)
    prod = v.double() * x.double()[ci.long()]
    ref = torch.zeros(m, device=rp.device, dtype=torch.float64).index_add_(0, rows, prod)
    mag = torch.zeros(m, device=rp.device, dtype=torch.float64).index_add_(0, rows, prod.abs_())
    return ref, mag


def _full_size_real_check(torch, rp, ci, v, x, dtype, tol, kernels=(1,)):
    from benchmark_spmv_using_csr5_b200 import handle as H
    m, n, nnz = rp.numel() - 1, x.numel(), ci.numel()
    ref, mag = _row_sums_fp64(torch, rp, ci, v, x)
    for kernel in kernels:
        h = H.anonymouslibHandle(m, n, dtype)
        assert h.inputCSR(nnz, rp, ci, v) == 0
        h.setX(x)
        h.setSigma(-1)
        h.set_option(H.OPT_KERNEL, _OPT_KERNEL[kernel])
        h.set_option(H.OPT_HOT_COLUMNS, _OPT_HOT[kernel])
        assert h.asCSR5() == 0
        y = torch.full((m,), float("nan"), device="cuda", dtype=dtype)
        assert h.spmv(1.0, y) == 0
        err = ((y.double() - ref).abs() / mag.clamp_min(1e-300)).max().item()
        assert err <= tol, f"kernel {kernel}: max row-wise relative error {err} > {tol}"
        assert h.destroy() == 0
        h.free()


def test_full_size_c2_real_valued(torch_cuda):
    """configs[1] with uniform (0,1] values (bench.py's inputs): every row within 1e-12 of the FP64 row sum
    (north_star bar: 1e-6)."""
    torch = torch_cuda
    m = 10_000_000
    rp, ci = M.device_banded(m, 16)
    v, x = M.device_values(ci.numel(), m, "real", torch.float64, "cuda")
    _full_size_real_check(torch, rp, ci, v, x, torch.float64, FP64_RTOL, kernels=(1, 5))


def test_full_size_c3_real_valued(torch_cuda):
    """configs[2] with uniform (0,1] values: hub rows of ~1e5 terms, carries through hundreds of tiles."""
    torch = torch_cuda
    rp, ci = M.device_rmat(22)
    n = rp.numel() - 1
    v, x = M.device_values(ci.numel(), n, "real", torch.float64, "cuda")
    _full_size_real_check(torch, rp, ci, v, x, torch.float64, FP64_RTOL, kernels=(1, 3))


def test_full_size_c5_rmat25_one_gpu(torch_cuda):
    """configs[4]'s matrix (R-MAT scale 25, ~5.2e8 nnz, 6.9 GB) on ONE GPU: bit-exact on integer-valued inputs
    against exact segment sums, and the asCSR5 -> asCSR round trip at that size (33.5 M rows, 57 % of them empty)."""
    torch = torch_cuda
    rp, ci = M.device_rmat(25)
    n = rp.numel() - 1
    v, x = M.device_values(ci.numel(), n, "int", torch.float64, "cuda")
    _full_size_check(torch, rp, ci, v, x, torch.float64, kernels=(1,))
    del v, x
    torch.cuda.empty_cache()
    v, x = M.device_values(ci.numel(), n, "real", torch.float64, "cuda")
    _full_size_real_check(torch, rp, ci, v, x, torch.float64, FP64_RTOL)


def test_cuda_y_vs_reference_avx2_directly(torch_cuda, oracle):
    """north_star: 'results match the reference AVX2 path's y within 1e-6 relative on FP64' -- the CUDA y against
    the y of the reference's own CSR5_avx2 (compiled from /root/reference into oracle/_ref), no oracle in between."""
    torch = torch_cuda
    if not oracle.ref_available():
        pytest.skip("oracle/_ref/libref_avx2.so not built (needs /root/reference at build time)")
    shapes = [("banded_300k", M.banded(300_000, 16)), ("rmat16", M.rmat(16)), ("example_c1", M.example_c1()),
              ("lap27_40", M.laplacian27(40)[0])]
    for name, A in shapes:
        for kind, tol in (("int", 0.0), ("real", 1e-6)):
            val, x = M.values(A.nnz, A.n, kind)
            y_avx2 = oracle.ref_avx2_spmv(A.m, A.n, A.row_ptr, A.col, val, x)
            h, _keep = _handle(torch, A, val, x, -1, kernel=1)
            y = _spmv(torch, h, A.m, torch.float64)
            if kind == "int":
                assert np.array_equal(y, y_avx2), name
            else:
                denom = np.maximum(np.abs(y_avx2), 1e-300)
                worst = float(np.max(np.abs(y - y_avx2) / denom))
                assert worst <= tol, f"{name}: {worst}"
                assert worst <= 1e-12, f"{name}: {worst} (expected ~1e-15)"
            h.free()


def test_sigma_rule_b200_is_bit_exact_too(torch_cuda, oracle):
    """CSR5B200_OPT_SIGMA_RULE: the reference's table stays the default (metadata word for word); the rule measured
    on B200 only changes which sigma AUTO picks -- both are bit-exact against the oracle run at the same sigma."""
    torch = torch_cuda
    from benchmark_spmv_using_csr5_b200 import handle as H
    for name, A, _sigma in CASES:
        if A.nnz == 0:
            continue
        for dt, tdt in ((np.float64, torch.float64), (np.float32, torch.float32)):
            val, x = M.values(A.nnz, A.n, "int", dt)
            y_ref = oracle.csr_spmv(A.m, A.row_ptr, A.col, val, x)
            for rule in (H.SIGMA_RULE_REFERENCE, H.SIGMA_RULE_B200):
                rp, ci, v, xd = _upload(torch, A, val, x)
                h = H.anonymouslibHandle(A.m, A.n, tdt)
                assert h.inputCSR(A.nnz, rp, ci, v) == 0
                h.setX(xd)
                assert h.set_option(H.OPT_SIGMA_RULE, rule) == 0
                h.setSigma(H.ANONYMOUSLIB_AUTO_TUNED_SIGMA)
                assert h.asCSR5() == 0
                s = h.info().sigma
                if rule == H.SIGMA_RULE_REFERENCE:
                    assert s == oracle.auto_sigma(A.m, A.nnz), name
                assert 4 <= s <= 32
                y = _spmv(torch, h, A.m, tdt)
                assert np.array_equal(y, y_ref), (name, rule, s)
                assert np.array_equal(y, oracle.csr5_spmv(A.m, A.n, A.row_ptr, A.col, val, x, s)), (name, rule, s)
                got, want = h.meta_to_host(), oracle.csr5_meta(A.m, A.nnz, s, A.row_ptr)
                assert np.array_equal(got["tile_ptr"], want.tile_ptr) and np.array_equal(got["desc"], want.desc)
                h.free()


def test_deterministic_carry_pass(torch_cuda, oracle):
    """CSR5B200_OPT_DETERMINISTIC: rows that span many tiles (their carries are otherwise added with floating-point
    atomics, in the order the warps retire) come out bit-identical from run to run, and still within tolerance."""
    torch = torch_cuda
    from benchmark_spmv_using_csr5_b200 import handle as H
    for name in ("hub_row_s12", "empty_rows_long_row_s4", "rmat12_auto", "m1_one_row", "example_c1_auto"):
        _n, A, sigma = [c for c in CASES if c[0] == name][0]
        for dt, tdt, tol in ((np.float64, torch.float64, FP64_RTOL), (np.float32, torch.float32, FP32_RTOL)):
            val, x = M.values(A.nnz, A.n, "real", dt)
            h, _keep = _handle(torch, A, val, x, sigma, kernel=1)
            assert h.set_option(H.OPT_DETERMINISTIC, 1) == 0
            ys = [_spmv(torch, h, A.m, tdt) for _ in range(6)]
            for y in ys[1:]:
                assert np.array_equal(y, ys[0]), f"{name}: run-to-run bits differ in deterministic mode"
            ref = oracle.csr_spmv_f32_acc64(A.m, A.row_ptr, A.col, val, x) if dt == np.float32 else \
                oracle.csr_spmv(A.m, A.row_ptr, A.col, val, x)
            assert np.allclose(ys[0], ref, rtol=tol, atol=1e-5 if dt == np.float32 else 0), name
            vali, xi = M.values(A.nnz, A.n, "int", dt)
            h2, _k2 = _handle(torch, A, vali, xi, sigma, kernel=1)
            h2.set_option(H.OPT_DETERMINISTIC, 1)
            assert np.array_equal(_spmv(torch, h2, A.m, tdt), oracle.csr_spmv(A.m, A.row_ptr, A.col, vali, xi)), name
            h.free()
            h2.free()


def test_unaligned_caller_arrays(torch_cuda, oracle):
    """col / val that are only element-aligned (views one element into a buffer): the 16-byte vector transpose and the
    bulk-TMA kernels need 16-byte alignment and must step aside for their scalar / direct-load forms."""
    torch = torch_cuda
    from benchmark_spmv_using_csr5_b200 import handle as H
    for name in ("example_c1_auto", "two_packet_s26_empty", "banded16_s16"):
        _n, A, sigma = [c for c in CASES if c[0] == name][0]
        for dt, tdt in ((np.float64, torch.float64), (np.float32, torch.float32)):
            val, x = M.values(A.nnz, A.n, "int", dt)
            big_c = torch.zeros(A.nnz + 1, device="cuda", dtype=torch.int32)
            big_v = torch.zeros(A.nnz + 1, device="cuda", dtype=tdt)
            ci, v = big_c[1:], big_v[1:]
            ci.copy_(torch.from_numpy(A.col))
            v.copy_(torch.from_numpy(val))
            assert ci.data_ptr() % 16 != 0
            rp, xd = torch.from_numpy(A.row_ptr).cuda(), torch.from_numpy(x).cuda()
            for kernel in (1, 2):
                h = H.anonymouslibHandle(A.m, A.n, tdt)
                assert h.inputCSR(A.nnz, rp, ci, v) == 0 and h.setX(xd) == 0
                h.setSigma(sigma)
                h.set_option(H.OPT_KERNEL, kernel)
                assert h.asCSR5() == 0
                s = h.info().sigma
                want = oracle.csr5_meta(A.m, A.nnz, s, A.row_ptr)
                assert np.array_equal(ci.cpu().numpy(), oracle.transpose(A.col, s, A.nnz, want.tile_ptr, True)), name
                assert np.array_equal(v.cpu().numpy(), oracle.transpose(val, s, A.nnz, want.tile_ptr, True)), name
                y = _spmv(torch, h, A.m, tdt)
                assert np.array_equal(y, oracle.csr_spmv(A.m, A.row_ptr, A.col, val, x)), name
                assert h.destroy() == 0
                assert np.array_equal(ci.cpu().numpy(), A.col) and np.array_equal(v.cpu().numpy(), val)
                h.free()


def test_carry_pass_is_skipped_when_no_tile_continues_a_row(torch_cuda, oracle):
    """16 nnz/row at sigma 16 (the headline configuration): every tile and the tail start on a row boundary, so spmv() is
    one launch; any matrix with a row crossing a tile boundary keeps the carry pass."""
    torch = torch_cuda
    for name, want_carries in (("banded16_s16", 0), ("rows_eq_tile_s4", 0), ("banded16_exact_multiple", 0),
                               ("hub_row_s12", 1), ("random_noempty_s8", 1), ("example_c1_auto", 1)):
        _n, A, sigma = [c for c in CASES if c[0] == name][0]
        val, x = M.values(A.nnz, A.n, "int")
        h, _keep = _handle(torch, A, val, x, sigma, kernel=1)
        y = _spmv(torch, h, A.m, torch.float64)
        i = h.info()
        assert i.has_carries == want_carries, name
        assert i.launches_per_spmv == (1 if not want_carries else 2) + (1 if i.needs_zero_fill else 0), name
        assert np.array_equal(y, oracle.csr_spmv(A.m, A.row_ptr, A.col, val, x)), name
        h.free()
