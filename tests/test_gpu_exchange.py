"""GPU tests (-m gpu) of what round 2 adds around the SpMV, all through the C ABI and all runnable on ONE GPU:

* y = alpha A x + beta y (csr5b200_spmv_axpby -- the reference's commented-out `beta`, anonymouslib_cuda.h:281)
  against the oracle's scalar statement;
* the overlapped all-gather step (csr5b200_spmv_allgather): row-block cut of the SpMV is bit-identical to the
  one-launch SpMV for every chunk count, every transport delivers every shard's rows to every shard;
* the single-process sharded handle (include/csr5_b200_sharded.h) with several shards on the same device -- the
  multi-GPU code path (peer copies, push kernels, in-kernel stores, flag / event barriers, double-buffered y,
  y -> x feedback) with device 0 standing in for every peer.
"""
import os

import numpy as np
import pytest

from benchmark_spmv_using_csr5_b200 import matrices as M
from tests.cases import small_cases

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASES = small_cases()
IDS = [c[0] for c in CASES]


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch


def _handle(torch, A, val, x, sigma):
    from benchmark_spmv_using_csr5_b200 import handle as H
    tdt = torch.float64 if val.dtype == np.float64 else torch.float32
    dev = "cuda"
    keep = (torch.from_numpy(A.row_ptr).to(dev), torch.from_numpy(A.col).to(dev), torch.from_numpy(val).to(dev),
            torch.from_numpy(x).to(dev))
    h = H.anonymouslibHandle(A.m, A.n, tdt)
    assert h.inputCSR(A.nnz, keep[0], keep[1], keep[2]) == 0
    assert h.setX(keep[3]) == 0
    h.setSigma(sigma)
    assert h.asCSR5() == 0
    return h, keep


# ---- beta ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dt", [np.float64, np.float32], ids=["f64", "f32"])
@pytest.mark.parametrize("name,A,sigma", CASES, ids=IDS)
def test_axpby_matches_oracle(torch_cuda, oracle, name, A, sigma, dt):
    torch = torch_cuda
    rng = np.random.default_rng(5)
    val, x = M.values(A.nnz, A.n, "int", dt)
    y0 = rng.integers(-9, 10, size=A.m).astype(dt)
    h, _keep = _handle(torch, A, val, x, sigma)
    for alpha, beta in ((1.0, 1.0), (2.0, -3.0), (-1.0, 0.5), (3.0, 0.0)):
        y = torch.from_numpy(y0.copy()).cuda()
        assert h.spmv_axpby(alpha, beta, y) == 0
        torch.cuda.synchronize()
        want = oracle.csr_axpby(A.m, A.row_ptr, A.col, val, x, alpha, beta, y0)
        assert np.array_equal(y.cpu().numpy(), want), f"{name}: alpha={alpha} beta={beta}"
    # real-valued: tolerance of the path (1e-12 rel of the row's magnitude in FP64, 2e-5 in FP32)
    val, x = M.values(A.nnz, A.n, "real", dt)
    h2, _keep2 = _handle(torch, A, val, x, sigma)
    y0r = rng.random(A.m).astype(dt)
    y = torch.from_numpy(y0r.copy()).cuda()
    assert h2.spmv_axpby(1.5, -0.25, y) == 0
    torch.cuda.synchronize()
    want = oracle.csr_axpby(A.m, A.row_ptr, A.col, val, x, 1.5, -0.25, y0r)
    scale = np.abs(oracle.csr_axpby(A.m, A.row_ptr, A.col, np.abs(val), np.abs(x), 1.5, 0.25, np.abs(y0r)))
    tol = 1e-12 if dt == np.float64 else 2e-5
    assert np.all(np.abs(y.cpu().numpy() - want) <= tol * np.maximum(scale, 1e-300)), name
    for hh in (h, h2):
        assert hh.destroy() == 0
        hh.free()


def test_axpby_beta_zero_never_reads_y(torch_cuda, oracle):
    torch = torch_cuda
    A = M.example_c1()
    val, x = M.values(A.nnz, A.n, "int")
    h, _keep = _handle(torch, A, val, x, -1)
    y = torch.full((A.m,), float("nan"), device="cuda", dtype=torch.float64)
    assert h.spmv_axpby(1.0, 0.0, y) == 0
    torch.cuda.synchronize()
    assert np.array_equal(y.cpu().numpy(), oracle.csr_spmv(A.m, A.row_ptr, A.col, val, x))
    h.free()


# ---- the row-block cut of the SpMV (world = 1: no peers, no barriers) ---------------------------------------
@pytest.mark.parametrize("name,A,sigma", CASES, ids=IDS)
def test_allgather_world1_chunks_bit_identical(torch_cuda, oracle, name, A, sigma):
    from benchmark_spmv_using_csr5_b200 import _lib
    torch = torch_cuda
    if A.nnz == 0:
        pytest.skip("empty matrix")
    for dt, tdt in ((np.float64, torch.float64), (np.float32, torch.float32)):
        val, x = M.values(A.nnz, A.n, "real", dt)   # real values: the cut must not change a single bit
        h, _keep = _handle(torch, A, val, x, sigma)
        y_one = torch.full((A.m,), float("nan"), device="cuda", dtype=tdt)
        assert h.spmv(1.0, y_one) == 0
        torch.cuda.synchronize()
        p = h.info().p
        # carries of multi-tile rows are added with atomics: only rows fed by more than one carry can differ
        # in the last bit between runs, so compare with the documented tolerance and bit-exactly on int data
        for chunks in (0, 1, 2, 3, 7, 64):
            y = torch.full((A.m,), float("nan"), device="cuda", dtype=tdt)
            ex = _lib.Csr5Exchange()
            ex.rank, ex.world = 0, 1
            ex.y_full[0] = y.data_ptr()
            ex.row_begin = 0
            ex.chunks = chunks
            assert h.spmv_allgather(1.0, 0.0, ex) == 0
            assert h.exchange_status() == 0
            assert np.allclose(y.cpu().numpy(), y_one.cpu().numpy(), rtol=1e-13 if dt == np.float64 else 1e-6, atol=0), \
                f"{name}: chunks={chunks} p={p}"
        vali, xi = M.values(A.nnz, A.n, "int", dt)
        h2, _k2 = _handle(torch, A, vali, xi, sigma)
        y = torch.full((A.m,), float("nan"), device="cuda", dtype=tdt)
        ex = _lib.Csr5Exchange()
        ex.rank, ex.world = 0, 1
        ex.y_full[0] = y.data_ptr()
        ex.chunks = 5
        assert h2.spmv_allgather(1.0, 0.0, ex) == 0
        assert h2.exchange_status() == 0
        assert np.array_equal(y.cpu().numpy(), oracle.csr_spmv(A.m, A.row_ptr, A.col, vali, xi)), name
        h.free()
        h2.free()


# ---- the single-process sharded handle, several shards on device 0 --------------------------------------------
TRANSPORTS = ["ce", "push", "inkernel"]


@pytest.mark.parametrize("transport", TRANSPORTS)
@pytest.mark.parametrize("name,A,sigma", CASES, ids=IDS)
def test_native_sharded_on_one_gpu(torch_cuda, oracle, name, A, sigma, transport):
    from benchmark_spmv_using_csr5_b200 import sharded as S
    if A.nnz == 0:
        pytest.skip("empty matrix")
    for dt in (np.float64, np.float32):
        val, x = M.values(A.nnz, A.n, "int", dt)
        y_ref = oracle.csr_spmv(A.m, A.row_ptr, A.col, val, x)
        for shards in (2, 3):
            sh = S.ShardedCsr5Native([0] * shards, dt)
            sh.inputCSR(A.m, A.n, A.row_ptr, A.col, val)
            assert np.array_equal(sh.bounds(), S.row_partition(A.row_ptr, shards)), "C++ and Python partition rules differ"
            sh.setSigma(sigma)
            sh.set_exchange(transport, chunks=3, push_ctas=4, barrier=S.BARRIER_EVENTS, timeout_ms=5000)
            sh.setX(x)
            sh.asCSR5()
            for _ in range(3):   # repeated steps alternate the two y buffers and stay exact
                sh.spmv(1.0)
            for g in range(shards):
                assert np.array_equal(sh.y(g), y_ref), f"{name} {dt.__name__} shards={shards} shard {g}"
            if transport != "inkernel":
                y_prev = sh.y(0)
                sh.spmv(2.0, -1.0)   # beta refers to the previous step's y
                for g in range(shards):
                    assert np.array_equal(sh.y(g), 2.0 * y_ref - y_prev), f"{name}: beta step, shard {g}"
            sh.destroy()


def test_native_sharded_iterate_feeds_y_back_as_x(torch_cuda, oracle):
    from benchmark_spmv_using_csr5_b200 import sharded as S
    A = M.banded(3000, 16)
    rng = np.random.default_rng(3)
    val = rng.integers(0, 3, size=A.nnz).astype(np.float64)   # small integers: three steps stay exact in FP64
    x = rng.integers(0, 3, size=A.n).astype(np.float64)
    want = x
    for _ in range(3):
        want = oracle.csr_spmv(A.m, A.row_ptr, A.col, val, want)
    for transport in ("ce", "push"):
        sh = S.ShardedCsr5Native([0, 0, 0], np.float64)
        sh.inputCSR(A.m, A.n, A.row_ptr, A.col, val)
        sh.set_exchange(transport, chunks=4, barrier=S.BARRIER_EVENTS, timeout_ms=5000)
        sh.setX(x)
        sh.asCSR5()
        sh.iterate(3)
        for g in range(3):
            assert np.array_equal(sh.y(g), want), transport
        # x is now the last iterate, held in a buffer of its own: further steps must not eat it
        want4 = oracle.csr_spmv(A.m, A.row_ptr, A.col, val, want)
        for _ in range(3):
            sh.spmv(1.0)
            for g in range(3):
                assert np.array_equal(sh.y(g), want4), transport
        sh.destroy()


def _rank0_of_two(torch, A, val, x, sigma, transport, flags0, flags1, y0, y1, timeout_ms=3000, chunks=3):
    """One spmv_allgather call as rank 0 of a world of 2 whose 'peer' lives in buffers on the same GPU."""
    from benchmark_spmv_using_csr5_b200 import _lib, handle as H
    from benchmark_spmv_using_csr5_b200 import sharded as S
    bounds = S.row_partition(A.row_ptr, 2)
    rp, ci, v = S.shard_csr(A.row_ptr, A.col, val, bounds[0], bounds[1])
    tdt = torch.float64 if val.dtype == np.float64 else torch.float32
    keep = [torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in (rp, ci, v, x)]
    h = H.anonymouslibHandle(int(bounds[1]), A.n, tdt)
    assert h.inputCSR(int(ci.size), keep[0], keep[1], keep[2]) == 0 and h.setX(keep[3]) == 0
    h.setSigma(sigma)
    assert h.asCSR5() == 0
    ex = _lib.Csr5Exchange()
    ex.rank, ex.world = 0, 2
    ex.y_full[0], ex.y_full[1] = y0.data_ptr(), y1.data_ptr()
    ex.flags[0], ex.flags[1] = flags0.data_ptr(), flags1.data_ptr()
    ex.row_begin, ex.chunks, ex.push_ctas, ex.timeout_ms = 0, chunks, 2, timeout_ms
    ex.transport = H.TRANSPORT_NAMES[transport]
    return h, ex, keep, int(bounds[1])


@pytest.mark.parametrize("transport", ["ce", "push", "inkernel"])
def test_flag_barrier_and_peer_delivery_as_rank0_of_two(torch_cuda, oracle, transport):
    """The multi-process step on one GPU, deterministically: this process is rank 0 of 2, the peer's y buffer and flag
    words are plain device buffers, and the peer's barrier arrivals are written in advance (epochs far ahead), so the
    device-side barrier never has to spin for a kernel that shares the GPU with it.  Checks: rank 0's rows land in
    the peer's buffer, both barrier slots signal the peer with increasing epochs, nothing times out."""
    torch = torch_cuda
    A = M.example_c1()
    val, x = M.values(A.nnz, A.n, "int", np.float64)
    y_ref = oracle.csr_spmv(A.m, A.row_ptr, A.col, val, x)
    flags0 = torch.zeros(64, device="cuda", dtype=torch.int32)
    flags1 = torch.zeros(64, device="cuda", dtype=torch.int32)
    flags0[1] = 1000      # slot 0 (entry), word [0 * world + peer]: the peer is far ahead
    flags0[2 + 1] = 1000  # slot 1 (exit)
    y0 = torch.full((A.m,), float("nan"), device="cuda", dtype=torch.float64)
    y1 = torch.full((A.m,), float("nan"), device="cuda", dtype=torch.float64)
    h, ex, _keep, rows0 = _rank0_of_two(torch, A, val, x, -1, transport, flags0, flags1, y0, y1)
    for step in range(1, 4):
        ex.entry_barrier = 1
        assert h.spmv_allgather(1.0, 0.0, ex) == 0
        assert h.exchange_status() == 0
        assert np.array_equal(y0[:rows0].cpu().numpy(), y_ref[:rows0])
        assert np.array_equal(y1[:rows0].cpu().numpy(), y_ref[:rows0]), "rank 0's rows did not reach the peer's buffer"
        assert torch.isnan(y1[rows0:]).all() and torch.isnan(y0[rows0:]).all()   # the peer's rows are the peer's job
        f1 = flags1.cpu().numpy()
        assert f1[0 * 2 + 0] == step and f1[1 * 2 + 0] == step, f1[:4]           # entry / exit arrival of rank 0
        y1[:rows0] = float("nan")
    h.free()


def test_flag_barrier_times_out_instead_of_hanging(torch_cuda):
    """A peer that never arrives: the barrier gives up after timeout_ms and the status call reports it."""
    torch = torch_cuda
    from benchmark_spmv_using_csr5_b200 import handle as H
    A = M.banded(2000, 16)
    val, x = M.values(A.nnz, A.n, "int", np.float64)
    flags0 = torch.zeros(64, device="cuda", dtype=torch.int32)
    flags1 = torch.zeros(64, device="cuda", dtype=torch.int32)
    y0 = torch.zeros(A.m, device="cuda", dtype=torch.float64)
    y1 = torch.zeros(A.m, device="cuda", dtype=torch.float64)
    h, ex, _keep, _rows0 = _rank0_of_two(torch, A, val, x, 16, "push", flags0, flags1, y0, y1, timeout_ms=200)
    assert h.spmv_allgather(1.0, 0.0, ex) == 0
    assert h.exchange_status() == H.EXCHANGE_TIMEOUT
    assert "timed out" in h.error_string(H.EXCHANGE_TIMEOUT)
    assert h.exchange_status() == 0   # reported once
    h.free()


def test_native_sharded_argument_errors(torch_cuda):
    from benchmark_spmv_using_csr5_b200 import _lib
    import ctypes as C
    lib = _lib.load_library()
    s = C.c_void_p()
    assert lib.csr5b200_sharded_create(0, (C.c_int * 1)(0), 8, C.byref(s)) == -101
    assert lib.csr5b200_sharded_create(1, (C.c_int * 1)(99), 8, C.byref(s)) == -101
    assert lib.csr5b200_sharded_create(1, (C.c_int * 1)(0), 2, C.byref(s)) == -5
    assert lib.csr5b200_sharded_create(1, (C.c_int * 1)(0), 8, C.byref(s)) == 0
    assert lib.csr5b200_sharded_spmv(s, 1.0, 0.0) == -101          # no matrix yet
    assert lib.csr5b200_sharded_set_exchange(s, 3, 0, 0, 0, 0) == -101   # multicast: multi-process binding only
    assert lib.csr5b200_sharded_destroy(s) == 0


@pytest.mark.parametrize("transport", [1, 2, 4])   # copy engine, push grid, in-kernel stores
def test_cpp_example_of_the_sharded_host_api(tmp_path, transport):
    """examples/sharded_spmv.cpp: the sharded SpMV driven from plain C++ (g++, no CUDA headers, no Python) through
    include/csr5_b200_sharded.h -- three shards on device 0, every shard's gathered y checked against the scalar loop."""
    import shutil
    import subprocess
    gxx = shutil.which("g++")
    if not gxx:
        pytest.skip("g++ not found")
    from benchmark_spmv_using_csr5_b200 import _lib
    exe = str(tmp_path / "sharded_spmv")
    libdir = os.path.dirname(_lib.LIB_PATH)
    r = subprocess.run([gxx, "-O2", "-std=c++17", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                        os.path.join(ROOT, "examples", "sharded_spmv.cpp"), "-L", libdir, "-lcsr5_b200",
                        f"-Wl,-rpath,{libdir}", "-o", exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([exe, "3", "200000", "5", str(transport), "1"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "Check... PASS!" in r.stdout, r.stdout + r.stderr


@pytest.mark.parametrize("scheme", [1, 2], ids=["fused", "push"])
@pytest.mark.parametrize("name,A,sigma", CASES, ids=IDS)
def test_legacy_scatter_two_local_destinations(torch_cuda, oracle, name, A, sigma, scheme):
    """csr5b200_spmv_scatter (the first-generation exchange, still shipped as transport IN_KERNEL / for A/B): the MULTI
    instantiations of the SpMV kernels with two destination buffers on the same GPU standing in for two GPUs."""
    import ctypes as C
    from benchmark_spmv_using_csr5_b200 import handle as H
    torch = torch_cuda
    if A.nnz == 0:
        pytest.skip("empty matrix")
    for dt, tdt in ((np.float64, torch.float64), (np.float32, torch.float32)):
        val, x = M.values(A.nnz, A.n, "int", dt)
        y_ref = oracle.csr_spmv(A.m, A.row_ptr, A.col, val, x)
        h, _keep = _handle(torch, A, val, x, sigma)
        assert h.set_option(H.OPT_EXCHANGE, scheme) == 0
        y_local = torch.full((A.m,), float("nan"), device="cuda", dtype=tdt)
        y_peer = torch.full((A.m,), float("nan"), device="cuda", dtype=tdt)
        dst = (C.c_void_p * 2)(y_local.data_ptr(), y_peer.data_ptr())
        for _ in range(2):
            assert h.spmv_scatter(1.0, y_local, dst, 2, False) == 0
        torch.cuda.synchronize()
        assert np.array_equal(y_local.cpu().numpy(), y_ref), f"{name}: local"
        assert np.array_equal(y_peer.cpu().numpy(), y_ref), f"{name}: peer"
        if scheme == 1 and A.m > 0:   # fused: the destination list must contain y_local (carries complete there)
            only_peer = (C.c_void_p * 1)(y_peer.data_ptr())
            assert h.spmv_scatter(1.0, y_local, only_peer, 1, False) == -101
        h.free()


def test_native_partition_by_row_cost_matches_python_rule(torch_cuda, oracle):
    """csr5b200_sharded_set_partition: the C++ host applies the same max(nnz, cost * rows) rule as sharded.row_partition,
    and the result stays exact whatever the partition."""
    from benchmark_spmv_using_csr5_b200 import sharded as S
    A = M.rmat(13)
    val, x = M.values(A.nnz, A.n, "int", np.float64)
    y_ref = oracle.csr_spmv(A.m, A.row_ptr, A.col, val, x)
    for cost in (0.0, 3.0, 8.0):
        sh = S.ShardedCsr5Native([0, 0, 0, 0], np.float64)
        sh.set_partition(cost)
        sh.inputCSR(A.m, A.n, A.row_ptr, A.col, val)
        assert np.array_equal(sh.bounds(), S.row_partition(A.row_ptr, 4, row_cost=cost)), cost
        sh.set_exchange("push", chunks=4, push_ctas=4, barrier=S.BARRIER_EVENTS, timeout_ms=5000)
        sh.setX(x)
        sh.asCSR5()
        sh.spmv(1.0)
        for g in range(4):
            assert np.array_equal(sh.y(g), y_ref), (cost, g)
        sh.destroy()
