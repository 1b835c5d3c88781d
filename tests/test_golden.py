"""Golden vectors: outputs of the reference's own CSR5_cuda backend recorded on a B200
(tests/golden/refcuda_*.npz, made by tests/golden/make_golden.py).

CPU part (not gpu): the oracle restatement reproduces them.  GPU part: so does the CUDA path."""
import glob
import os

import numpy as np
import pytest

from benchmark_spmv_using_csr5_b200 import matrices as M
from tests.cases import small_cases
from tests.golden.make_golden import digest

HERE = os.path.dirname(os.path.abspath(__file__))
FIXTURES = sorted(glob.glob(os.path.join(HERE, "golden", "refcuda_*.npz")))
BY_NAME = {c[0]: c for c in small_cases()}


def _load(path):
    base = os.path.basename(path)[len("refcuda_"):-len(".npz")]
    name, tag = base.rsplit("_", 1)
    dt = np.float64 if tag == "f64" else np.float32
    _, A, sigma = BY_NAME[name]
    g = np.load(path)
    vi, xi = M.values(A.nnz, A.n, "int", dt)
    assert str(g["input_digest"]) == digest(A.row_ptr) + digest(A.col) + digest(vi) + digest(xi), \
        "seeded generators no longer reproduce the fixture's inputs"
    return name, A, sigma, dt, g, vi, xi


def test_fixtures_present():
    assert len(FIXTURES) >= 20, "golden fixtures missing: run tests/golden/make_golden.py on a GPU box"


@pytest.mark.parametrize("path", FIXTURES, ids=[os.path.basename(p)[8:-4] for p in FIXTURES])
def test_oracle_reproduces_reference_cuda_golden(oracle, path):
    name, A, sigma, dt, g, vi, xi = _load(path)
    sc = g["scalars"]
    s, by, bs, npk, p, noff, tail = (int(v) for v in sc)
    meta = oracle.csr5_meta(A.m, A.nnz, s, A.row_ptr)
    assert (meta.sigma, meta.bit_y, meta.bit_ss, meta.num_packet, meta.p, meta.tail_start) == (s, by, bs, npk, p, tail)
    assert np.array_equal(meta.tile_ptr, g["tile_ptr"])
    live = (p - 1) * 32 * npk
    assert np.array_equal(meta.desc[:live], g["desc"][:live])
    if (g["tile_ptr"][:max(p - 1, 0)] >> 31).any():
        # the reference's rounded-up grid also counts the tail tile's segments into num_offsets
        # (SURVEY.md App. B); only the entries below desc_off_ptr[p - 1] are ever read
        assert np.array_equal(meta.desc_off_ptr[:p], g["desc_off_ptr"][:p])
        n = int(g["desc_off_ptr"][p - 1])
        assert n <= meta.num_offsets <= noff
        assert np.array_equal(meta.desc_off[:n], g["desc_off"][:n])
    assert digest(oracle.transpose(A.col, s, A.nnz, meta.tile_ptr, True)) == str(g["col5_digest"])
    assert digest(oracle.transpose(vi, s, A.nnz, meta.tile_ptr, True)) == str(g["val5_digest"])
    assert np.array_equal(oracle.csr5_spmv(A.m, A.n, A.row_ptr, A.col, vi, xi, sigma), g["y_int"])
    vr, xr = M.values(A.nnz, A.n, "real", dt)
    y = oracle.csr5_spmv(A.m, A.n, A.row_ptr, A.col, vr, xr, sigma)
    if dt == np.float64:
        assert np.allclose(y, g["y_real"], rtol=1e-12, atol=0)
    else:
        assert np.allclose(y, g["y_real"], rtol=2e-5, atol=1e-5)


@pytest.mark.gpu
@pytest.mark.parametrize("path", FIXTURES, ids=[os.path.basename(p)[8:-4] for p in FIXTURES])
def test_cuda_path_reproduces_reference_cuda_golden(path):
    import torch
    from benchmark_spmv_using_csr5_b200 import handle as H
    name, A, sigma, dt, g, vi, xi = _load(path)
    tdt = torch.float64 if dt == np.float64 else torch.float32
    for kind, want in (("int", g["y_int"]), ("real", g["y_real"])):
        val, x = M.values(A.nnz, A.n, kind, dt)
        rp, ci = torch.from_numpy(A.row_ptr).cuda(), torch.from_numpy(A.col).cuda()
        v, xd = torch.from_numpy(val).cuda(), torch.from_numpy(x).cuda()
        h = H.anonymouslibHandle(A.m, A.n, tdt)
        assert h.inputCSR(A.nnz, rp, ci, v) == 0 and h.setX(xd) == 0
        h.setSigma(sigma)
        h.set_option(H.OPT_HOT_COLUMNS, 0)   # compare col5 with the reference's: no column tagging
        assert h.asCSR5() == 0
        y = torch.full((A.m,), float("nan"), device="cuda", dtype=tdt)
        assert h.spmv(1.0, y) == 0
        torch.cuda.synchronize()
        y = y.cpu().numpy()
        if kind == "int":
            assert np.array_equal(y, want), name
            meta = h.meta_to_host()
            assert np.array_equal(meta["tile_ptr"], g["tile_ptr"])
            live = (meta["p"] - 1) * 32 * meta["num_packet"]
            assert np.array_equal(meta["desc"][:live], g["desc"][:live])
            assert digest(ci.cpu().numpy()) == str(g["col5_digest"])
            assert digest(v.cpu().numpy()) == str(g["val5_digest"])
        elif dt == np.float64:
            assert np.allclose(y, want, rtol=1e-12, atol=0), name
        else:
            assert np.allclose(y, want, rtol=2e-5, atol=1e-5), name
        h.free()
