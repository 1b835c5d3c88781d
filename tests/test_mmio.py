"""Matrix-Market ingest: the reader reproduces the reference loader's CSR
(CSR5_cuda/main.cu:157-312) -- checked against a literal loop restatement of that loader on small
files and against scipy.io.mmread for the matrix it represents."""
import io

import numpy as np
import pytest
import scipy.io
import scipy.sparse as sp

from benchmark_spmv_using_csr5_b200 import matrices as M
from benchmark_spmv_using_csr5_b200 import mmio


def _reference_loader(text):
    """Loop restatement of main.cu:211-306 (fscanf per entry, symmetric expansion, counting sort)."""
    lines = [ln for ln in text.splitlines() if ln.strip()]
    banner = lines[0].lower().split()
    field, sym = banner[3], banner[4]
    body = [ln for ln in lines[1:] if not ln.startswith("%")]
    m, n, nz = (int(t) for t in body[0].split())
    ent = []
    for ln in body[1:1 + nz]:
        t = ln.split()
        ent.append((int(t[0]) - 1, int(t[1]) - 1, 1.0 if field == "pattern" else float(t[2])))
    symmetric = sym in ("symmetric", "hermitian")
    cnt = [0] * (m + 1)
    for r, c, _ in ent:
        cnt[r] += 1
        if symmetric and r != c:
            cnt[c] += 1
    rp = [0] * (m + 1)
    for i in range(m):
        rp[i + 1] = rp[i] + cnt[i]
    fill = [0] * m
    col = [0] * rp[m]
    val = [0.0] * rp[m]
    for r, c, v in ent:
        col[rp[r] + fill[r]] = c; val[rp[r] + fill[r]] = v; fill[r] += 1
        if symmetric and r != c:
            col[rp[c] + fill[c]] = r; val[rp[c] + fill[c]] = v; fill[c] += 1
    return m, n, np.array(rp, np.int32), np.array(col, np.int32), np.array(val)


FILES = {
    "general_real": "%%MatrixMarket matrix coordinate real general\n% c\n4 5 6\n3 1 1.5\n1 2 2\n3 5 -1\n1 1 4\n4 4 9\n3 1 7\n",
    "symmetric_real": "%%MatrixMarket matrix coordinate real symmetric\n4 4 5\n1 1 1\n3 1 2\n4 2 3\n4 4 5\n3 2 6\n",
    "pattern_general": "%%MatrixMarket matrix coordinate pattern general\n3 3 4\n1 3\n2 2\n3 1\n1 1\n",
    "integer_symmetric": "%%MatrixMarket matrix coordinate integer symmetric\n3 3 3\n2 1 4\n3 3 7\n3 1 -2\n",
    "hermitian_real": "%%MatrixMarket matrix coordinate real hermitian\n3 3 2\n2 1 4\n3 3 7\n",
    "skew": "%%MatrixMarket matrix coordinate real skew-symmetric\n3 3 2\n2 1 4\n3 2 7\n",
}


@pytest.mark.parametrize("name", sorted(FILES))
def test_reader_matches_reference_loader(name):
    text = FILES[name]
    m, n, rp, col, val = mmio.read_mtx(io.StringIO(text))
    rm, rn, rrp, rcol, rval = _reference_loader(text)
    assert (m, n) == (rm, rn)
    assert np.array_equal(rp, rrp) and np.array_equal(col, rcol) and np.array_equal(val, rval)
    if name != "skew":  # the reference does not expand skew-symmetric files; scipy does
        want = scipy.io.mmread(io.StringIO(text), spmatrix=False).toarray()
        got = sp.csr_matrix((val, col, rp), shape=(m, n)).toarray()   # duplicates are summed by both
        assert np.array_equal(got, want)


def test_round_trip_and_rejections(tmp_path):
    A = M.example_c1()
    val, _ = M.values(A.nnz, A.n, "real")
    p = str(tmp_path / "a.mtx")
    mmio.write_mtx(p, A.m, A.n, A.row_ptr, A.col, val)
    m, n, rp, col, v = mmio.read_mtx(p)
    assert (m, n) == (A.m, A.n) and np.array_equal(rp, A.row_ptr) and np.array_equal(col, A.col)
    assert np.array_equal(v, val)
    for bad in ("%%MatrixMarket matrix coordinate complex general\n1 1 1\n1 1 1 0\n",
                "%%MatrixMarket matrix array real general\n1 1\n1\n", "garbage\n"):
        with pytest.raises(mmio.MatrixMarketError):
            mmio.read_mtx(io.StringIO(bad))


@pytest.mark.gpu
@pytest.mark.parametrize("vt", ["double", "float"])
def test_cli_parity_run(tmp_path, capsys, vt):
    """`./spmv file.mtx` equivalent: same report lines, self-check passes."""
    from benchmark_spmv_using_csr5_b200 import cli
    A = M.example_c1()
    p = str(tmp_path / "example.mtx")
    mmio.write_mtx(p, A.m, A.n, A.row_ptr, A.col, None, field="pattern")
    csv = str(tmp_path / "results.csv")
    rc = cli.main([p, "--value-type", vt, "--num-run", "20", "--seed", "1", "--results-csv", csv])
    out = capsys.readouterr().out
    assert rc == 0 and "Check... PASS!" in out
    assert f" ( {A.m}, {A.n} ) nnz = {A.nnz}" in out
    assert "CSR->CSR5 time = " in out and "CSR5-based SpMV time = " in out and "omega = 32, sigma = " in out
    assert open(csv).read().startswith(p + ",")


def test_empty_file_is_a_valid_matrix():
    """A coordinate file with nnz = 0 (the reference loader accepts it: its loops simply do not run)."""
    m, n, rp, col, val = mmio.read_mtx(io.StringIO("%%MatrixMarket matrix coordinate real general\n% nothing\n3 4 0\n"))
    assert (m, n) == (3, 4) and rp.tolist() == [0, 0, 0, 0] and col.size == 0 and val.size == 0
    with pytest.raises(mmio.MatrixMarketError):
        mmio.read_mtx(io.StringIO("%%MatrixMarket matrix coordinate real general\n3 4 2\n1 1 1.0\n"))


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(FILES))
def test_device_coo_to_csr_matches_reference_loader(name):
    """csr5b200_coo_to_csr (symmetric expansion + stable radix sort by row on the GPU) against the loop restatement
    of the reference's loader, entry for entry: order inside a row, duplicates, mirrored entries."""
    text = FILES[name]
    for dt in (np.float64, np.float32):
        m, n, rp, col, val = mmio.read_mtx_device(io.StringIO(text), dt)
        rm, rn, rrp, rcol, rval = _reference_loader(text)
        assert (m, n) == (rm, rn)
        assert np.array_equal(rp.cpu().numpy(), rrp) and np.array_equal(col.cpu().numpy(), rcol)
        assert np.array_equal(val.cpu().numpy(), rval.astype(dt))


@pytest.mark.gpu
def test_device_coo_to_csr_large_shuffled_symmetric():
    """Many blocks, several radix passes (m > 2^16), long rows, duplicates, shuffled file order, mirrored entries:
    the device result equals the host restatement of the loader (mmio.read_mtx is pinned to it above)."""
    import torch
    rng = np.random.default_rng(11)
    m = n = 70_000
    nnz = 600_000
    r = rng.integers(0, m, nnz).astype(np.int32)
    c = rng.integers(0, n, nnz).astype(np.int32)
    r[:50_000] = 12345                       # a row of 50 000 entries spread over the file
    c[100:200] = r[100:200]                  # diagonal entries are not mirrored
    perm = rng.permutation(nnz)
    r, c = r[perm], c[perm]
    v = rng.integers(1, 100, nnz).astype(np.float64)
    for symmetric in (False, True):
        rp, col, val = mmio.coo_to_csr_device(m, n, torch.from_numpy(r).cuda(), torch.from_numpy(c).cuda(),
                                              torch.from_numpy(v).cuda(), symmetric)
        rr, cc, vv = r.astype(np.int64), c.astype(np.int64), v
        if symmetric:
            off = rr != cc
            keep = np.stack([np.ones_like(off), off], 1).reshape(-1)
            rr, cc, vv = (np.stack([rr, cc], 1).reshape(-1)[keep], np.stack([cc, rr], 1).reshape(-1)[keep],
                          np.stack([vv, vv], 1).reshape(-1)[keep])
        order = np.argsort(rr, kind="stable")
        want_rp = np.concatenate([[0], np.cumsum(np.bincount(rr, minlength=m))]).astype(np.int32)
        assert np.array_equal(rp.cpu().numpy(), want_rp)
        assert np.array_equal(col.cpu().numpy(), cc[order].astype(np.int32))
        assert np.array_equal(val.cpu().numpy(), vv[order])
    # pattern (no values) and out-of-range indices
    rp, col, val = mmio.coo_to_csr_device(m, n, torch.from_numpy(r).cuda(), torch.from_numpy(c).cuda(), None, False)
    assert float(val.sum()) == nnz
    bad = r.copy()
    bad[7] = m
    with pytest.raises(mmio.MatrixMarketError):
        mmio.coo_to_csr_device(m, n, torch.from_numpy(bad).cuda(), torch.from_numpy(c).cuda(), None, False)


@pytest.mark.gpu
def test_device_ingest_feeds_the_spmv(tmp_path, oracle):
    """File -> device COO -> device CSR -> CSR5 -> y, against the oracle on the host-loaded CSR of the same file."""
    import torch
    from benchmark_spmv_using_csr5_b200 import handle as H
    A = M.example_c1()
    val, x = M.values(A.nnz, A.n, "int")
    p = str(tmp_path / "example.mtx")
    mmio.write_mtx(p, A.m, A.n, A.row_ptr, A.col, val)
    m, n, rp, col, v = mmio.read_mtx_device(p)
    h = H.anonymouslibHandle(m, n, torch.float64)
    assert h.inputCSR(int(col.numel()), rp, col, v) == 0 and h.setX(torch.from_numpy(x).cuda()) == 0
    h.setSigma(-1)
    assert h.asCSR5() == 0
    y = torch.empty(m, device="cuda", dtype=torch.float64)
    assert h.spmv(1.0, y) == 0
    torch.cuda.synchronize()
    assert np.array_equal(y.cpu().numpy(), oracle.csr_spmv(A.m, A.row_ptr, A.col, val, x))
    h.free()
