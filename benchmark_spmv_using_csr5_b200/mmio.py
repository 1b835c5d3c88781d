"""Matrix-Market coordinate files -> CSR, with the semantics of the reference's loader
(CSR5_cuda/main.cu:157-312; banner/size rules of its vendored NIST reader, mmio.h:254-367):

* ``matrix coordinate {real|integer|pattern} {general|symmetric|hermitian|skew-symmetric}``;
  ``complex`` and ``array`` are rejected (main.cu:176-180);
* indices 1-based in the file, 0-based here; pattern entries get the value 1.0;
* ``symmetric`` and ``hermitian`` files are expanded: every off-diagonal (i, j) also yields (j, i),
  emitted right after it; ``skew-symmetric`` is NOT expanded (the reference only tests
  mm_is_symmetric / mm_is_hermitian, main.cu:192);
* COO -> CSR by a stable counting sort on the row: within a row the entries keep file order, columns
  are NOT sorted, duplicates are kept (main.cu:264-306).

The reference then discards the file's values and draws ``rand() % 10`` (main.cu:314-326);
``reference_values`` does the same with a seedable generator.
"""
from __future__ import annotations

import io

import numpy as np


class MatrixMarketError(ValueError):
    pass


def _parse_header(f):
    banner = f.readline()
    tok = banner.strip().split()
    if len(tok) != 5 or tok[0] != "%%MatrixMarket":
        raise MatrixMarketError("Could not process Matrix Market banner.")
    _, obj, fmt, field, sym = (t.lower() for t in tok)
    if obj != "matrix" or fmt != "coordinate":
        raise MatrixMarketError("only 'matrix coordinate' files are supported")
    if field == "complex":
        raise MatrixMarketError("Sorry, data type 'COMPLEX' is not supported.")
    if field not in ("real", "integer", "pattern", "double"):
        raise MatrixMarketError(f"unknown field '{field}'")
    if sym not in ("general", "symmetric", "hermitian", "skew-symmetric"):
        raise MatrixMarketError(f"unknown symmetry '{sym}'")
    line = f.readline()
    while line and (line.startswith("%") or not line.strip()):
        line = f.readline()
    try:
        m, n, nnz = (int(t) for t in line.split()[:3])
    except Exception as e:
        raise MatrixMarketError("could not read the size line") from e
    return field, sym, m, n, nnz


def read_coo(path_or_file, dtype=np.float64):
    """The file's entries as they stand: (field, symmetry, m, n, rows int32, cols int32, vals dtype | None for
    pattern files), 0-based, file order.  Text parsing is the host's part of the ingest (main.cu:211-237)."""
    f = open(path_or_file, "r") if isinstance(path_or_file, (str, bytes)) else path_or_file
    try:
        field, sym, m, n, nnz_file = _parse_header(f)
        ncol = 2 if field == "pattern" else 3
        body = f.read()
    finally:
        if isinstance(path_or_file, (str, bytes)):
            f.close()
    if nnz_file == 0 or not body.strip():
        arr = np.empty((0, ncol))   # a valid file with no entries (the reference loader accepts it)
    else:
        try:
            import pandas as pd
            try:
                df = pd.read_csv(io.StringIO(body), sep=r"\s+", header=None, comment="%", nrows=nnz_file,
                                 usecols=range(ncol), engine="c", float_precision="round_trip")
            except (pd.errors.EmptyDataError, pd.errors.ParserError, ValueError) as e:
                raise MatrixMarketError(f"could not parse the entries: {e}") from e
            arr = df.to_numpy(dtype=np.float64)
        except ImportError:  # pragma: no cover
            arr = np.loadtxt(io.StringIO(body), comments="%", max_rows=nnz_file, usecols=range(ncol), ndmin=2)
    if arr.shape[0] != nnz_file:
        raise MatrixMarketError(f"expected {nnz_file} entries, found {arr.shape[0]}")
    r = arr[:, 0].astype(np.int64) - 1
    c = arr[:, 1].astype(np.int64) - 1
    v = arr[:, 2].astype(dtype) if ncol == 3 else None
    if nnz_file and (r.min() < 0 or r.max() >= m or c.min() < 0 or c.max() >= n):
        raise MatrixMarketError("index out of range")
    return field, sym, m, n, r.astype(np.int32), c.astype(np.int32), v


def read_mtx(path_or_file, dtype=np.float64):
    """-> (m, n, row_ptr int32 (m+1), col int32 (nnz), val dtype (nnz)) in the reference's CSR order (host numpy
    restatement of main.cu:239-306; the device path is ``read_mtx_device``)."""
    _field, sym, m, n, r, c, v = read_coo(path_or_file, dtype)
    r, c = r.astype(np.int64), c.astype(np.int64)
    if v is None:
        v = np.ones(len(r), dtype)
    if sym in ("symmetric", "hermitian"):
        # interleave (i, j) and, for off-diagonal entries, (j, i): the reference's emission order
        off = r != c
        rr = np.stack([r, c], 1).reshape(-1)
        cc = np.stack([c, r], 1).reshape(-1)
        vv = np.stack([v, v], 1).reshape(-1)
        keep = np.stack([np.ones_like(off), off], 1).reshape(-1)
        r, c, v = rr[keep], cc[keep], vv[keep]
    order = np.argsort(r, kind="stable")
    row_ptr = np.zeros(m + 1, np.int64)
    np.add.at(row_ptr, r + 1, 1)
    row_ptr = np.cumsum(row_ptr)
    if row_ptr[-1] >= 2 ** 31:
        raise MatrixMarketError("nnz does not fit 32-bit indices")
    return m, n, row_ptr.astype(np.int32), c[order].astype(np.int32), np.ascontiguousarray(v[order])


def coo_to_csr_device(m, n, rows, cols, vals, symmetric: bool, dtype=None):
    """COO -> CSR on the GPU (csr5b200_coo_to_csr: symmetric expansion + stable sort by row as CUDA kernels,
    main.cu:239-306).  rows / cols: int32 CUDA tensors in file order; vals: CUDA tensor or None (pattern: all 1).
    Returns (row_ptr int32 (m+1), col int32, val) CUDA tensors."""
    import ctypes as C

    import torch

    from . import _lib
    lib = _lib.load_library()
    dtype = (vals.dtype if vals is not None else torch.float64) if dtype is None else dtype
    vb = 8 if dtype == torch.float64 else 4
    dev = rows.device
    nnz = int(rows.numel())
    cap = nnz * (2 if symmetric else 1)
    row_ptr = torch.empty(m + 1, device=dev, dtype=torch.int32)
    col = torch.empty(max(cap, 1), device=dev, dtype=torch.int32)
    val = torch.empty(max(cap, 1), device=dev, dtype=dtype)
    out = C.c_int(0)
    stream = torch.cuda.current_stream(dev).cuda_stream
    p = lambda t: C.c_void_p(t.data_ptr()) if t is not None and t.numel() else C.c_void_p(0)  # noqa: E731
    err = lib.csr5b200_coo_to_csr(int(m), int(n), nnz, p(rows), p(cols), p(vals), vb, 1 if symmetric else 0,
                                  p(row_ptr), C.c_void_p(col.data_ptr()), C.c_void_p(val.data_ptr()), cap,
                                  C.byref(out), C.c_void_p(stream))
    if err:
        raise MatrixMarketError(f"csr5b200_coo_to_csr: {lib.csr5b200_error_string(err).decode()}")
    return row_ptr, col[:out.value], val[:out.value]


def read_mtx_device(path_or_file, dtype=np.float64, device="cuda"):
    """Matrix-Market file -> CSR in HBM: the text is parsed on the host, the triples are uploaded as they stand, and
    the reference loader's COO -> CSR (main.cu:239-306) runs as CUDA kernels.  Returns (m, n, row_ptr, col, val)
    with CUDA tensors."""
    import torch
    _field, sym, m, n, r, c, v = read_coo(path_or_file, dtype)
    tdt = torch.float64 if np.dtype(dtype) == np.float64 else torch.float32
    rows, cols = torch.from_numpy(r).to(device), torch.from_numpy(c).to(device)
    vals = torch.from_numpy(np.ascontiguousarray(v)).to(device) if v is not None else None
    rp, col, val = coo_to_csr_device(m, n, rows, cols, vals, sym in ("symmetric", "hermitian"), tdt)
    return m, n, rp, col, val


def write_mtx(path, m, n, row_ptr, col, val=None, field="real", symmetry="general"):
    """Writes the CSR entries in row order as a coordinate file (no symmetry folding is applied: with
    symmetry != 'general' the caller passes the triangle it wants stored)."""
    rows = np.repeat(np.arange(m, dtype=np.int64), np.diff(np.asarray(row_ptr, np.int64)))
    with open(path, "w") as f:
        f.write(f"%%MatrixMarket matrix coordinate {field} {symmetry}\n% written by benchmark_spmv_using_csr5_b200\n")
        f.write(f"{m} {n} {len(col)}\n")
        if field == "pattern":
            np.savetxt(f, np.column_stack([rows + 1, np.asarray(col, np.int64) + 1]), fmt="%d %d")
        elif field == "integer":
            np.savetxt(f, np.column_stack([rows + 1, np.asarray(col, np.int64) + 1, np.asarray(val).astype(np.int64)]),
                       fmt="%d %d %d")
        else:
            out = np.column_stack([rows + 1, np.asarray(col, np.int64) + 1, np.asarray(val, np.float64)])
            np.savetxt(f, out, fmt="%d %d %.17g")


def reference_values(nnz, n, dtype=np.float64, seed=None):
    """val, x drawn from {0..9} as the reference does with rand() % 10 (main.cu:314-326).  The
    reference seeds with time(NULL); pass a seed for reproducible runs."""
    rng = np.random.default_rng(seed)
    return rng.integers(0, 10, nnz).astype(dtype), rng.integers(0, 10, n).astype(dtype)
