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


def read_mtx(path_or_file, dtype=np.float64):
    """-> (m, n, row_ptr int32 (m+1), col int32 (nnz), val dtype (nnz)) in the reference's CSR order."""
    f = open(path_or_file, "r") if isinstance(path_or_file, (str, bytes)) else path_or_file
    try:
        field, sym, m, n, nnz_file = _parse_header(f)
        ncol = 2 if field == "pattern" else 3
        body = f.read()
    finally:
        if isinstance(path_or_file, (str, bytes)):
            f.close()
    try:
        import pandas as pd
        df = pd.read_csv(io.StringIO(body), sep=r"\s+", header=None, comment="%", nrows=nnz_file,
                         usecols=range(ncol), engine="c",
                         float_precision="round_trip")
        arr = df.to_numpy(dtype=np.float64)
    except ImportError:  # pragma: no cover
        arr = np.loadtxt(io.StringIO(body), comments="%", max_rows=nnz_file, usecols=range(ncol), ndmin=2)
    if arr.shape[0] != nnz_file:
        raise MatrixMarketError(f"expected {nnz_file} entries, found {arr.shape[0]}")
    r = arr[:, 0].astype(np.int64) - 1
    c = arr[:, 1].astype(np.int64) - 1
    v = arr[:, 2].astype(dtype) if ncol == 3 else np.ones(nnz_file, dtype)
    if nnz_file and (r.min() < 0 or r.max() >= m or c.min() < 0 or c.max() >= n):
        raise MatrixMarketError("index out of range")
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
