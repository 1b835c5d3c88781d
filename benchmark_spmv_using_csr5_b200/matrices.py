"""Synthetic CSR inputs for the configurations named in BASELINE.json / SURVEY.md section 8(d).

Host (numpy) generators are used by the tests and the CPU baseline; device (torch) generators
build the full-size benchmark matrices directly in HBM.  All indices are int32, 0-based.

Value passes:
  * ``"int"``  -- val, x drawn from {0..9}: the reference's own input distribution
    (CSR5_cuda/main.cu:314-326).  Every partial sum is exact in FP64/FP32, so any summation
    order is bit-identical -- the regime in which "bit-exact" parity is defined.
  * ``"real"`` -- uniform (0, 1].
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


@dataclass
class HostCsr:
    m: int
    n: int
    row_ptr: np.ndarray  # (m+1,) int32
    col: np.ndarray      # (nnz,) int32
    name: str = ""

    @property
    def nnz(self) -> int:
        return int(self.row_ptr[-1])


# ------------------------------------------------------------------------------------------------
# host generators
# ------------------------------------------------------------------------------------------------

def banded(m: int, per_row: int = 16) -> HostCsr:
    """C2 shape: row i holds columns (i - per_row/2 ... i + per_row/2 - 1) mod n, in that order."""
    half = per_row // 2
    row_ptr = (np.arange(m + 1, dtype=np.int64) * per_row).astype(np.int32)
    col = (np.arange(m, dtype=np.int64)[:, None] + np.arange(-half, per_row - half)[None, :]) % m
    return HostCsr(m, m, row_ptr, col.reshape(-1).astype(np.int32), f"banded{per_row}_{m}")


def rmat(scale: int, edge_factor: int = 16, seed: int | None = None,
         abc=(0.57, 0.19, 0.19)) -> HostCsr:
    """C3/C5 shape: R-MAT, bits drawn LSB first, duplicates merged, self loops kept, sorted by
    (row, col), no vertex permutation."""
    a, b, c = abc
    n = 1 << scale
    ne = edge_factor * n
    rng = np.random.default_rng(scale if seed is None else seed)
    row = np.zeros(ne, np.int64)
    colv = np.zeros(ne, np.int64)
    for bit in range(scale):
        u = rng.random(ne)
        row |= (u >= a + b).astype(np.int64) << bit
        colv |= (((u >= a) & (u < a + b)) | (u >= a + b + c)).astype(np.int64) << bit
    key = np.unique(row * n + colv)
    row = key // n
    colv = key % n
    row_ptr = np.zeros(n + 1, np.int64)
    np.add.at(row_ptr, row + 1, 1)
    row_ptr = np.cumsum(row_ptr)
    return HostCsr(n, n, row_ptr.astype(np.int32), colv.astype(np.int32), f"rmat{scale}")


def laplacian27(nx: int, ny: int | None = None, nz: int | None = None):
    """C4 shape: 27-point stencil on an nx*ny*nz grid, x fastest, truncated at the boundaries,
    columns ascending.  Returns (HostCsr, val) with val = 26 on the diagonal, -1 elsewhere."""
    ny = nx if ny is None else ny
    nz = nx if nz is None else nz
    m = nx * ny * nz
    idx = np.arange(m, dtype=np.int64)
    ix, iy, iz = idx % nx, (idx // nx) % ny, idx // (nx * ny)
    cols, valid = [], []
    for dz in (-1, 0, 1):
        for dy in (-1, 0, 1):
            for dx in (-1, 0, 1):
                ok = ((ix + dx >= 0) & (ix + dx < nx) & (iy + dy >= 0) & (iy + dy < ny) &
                      (iz + dz >= 0) & (iz + dz < nz))
                cols.append(idx + dx + dy * nx + dz * nx * ny)
                valid.append(ok)
    cols = np.stack(cols, 1)
    valid = np.stack(valid, 1)
    row_ptr = np.concatenate([[0], np.cumsum(valid.sum(1))])
    col = cols[valid]
    val = np.where(col == np.repeat(idx, valid.sum(1)), 26.0, -1.0)
    return HostCsr(m, m, row_ptr.astype(np.int32), col.astype(np.int32), f"lap27_{nx}x{ny}x{nz}"), val


def example_c1(seed: int = 2015) -> HostCsr:
    """Stand-in for the reference's missing ``example.mtx`` (README.md:27):
    m = n = 10000, row i empty if i % 13 == 5, row 1234 has 5000 nnz, others 1..24 nnz,
    columns uniform (unsorted, duplicates possible, like the reference's COO->CSR)."""
    m = n = 10000
    rng = np.random.default_rng(seed)
    cnt = rng.integers(1, 25, size=m)
    cnt[np.arange(m) % 13 == 5] = 0
    cnt[1234] = 5000
    row_ptr = np.concatenate([[0], np.cumsum(cnt)])
    col = rng.integers(0, n, size=int(row_ptr[-1]))
    return HostCsr(m, n, row_ptr.astype(np.int32), col.astype(np.int32), "example_c1")


def from_row_counts(counts, n: int, seed: int = 0, name: str = "custom") -> HostCsr:
    """Arbitrary row-length profile with uniform random columns (adversarial test shapes)."""
    counts = np.asarray(counts, np.int64)
    rng = np.random.default_rng(seed)
    row_ptr = np.concatenate([[0], np.cumsum(counts)])
    col = rng.integers(0, n, size=int(row_ptr[-1]))
    return HostCsr(len(counts), n, row_ptr.astype(np.int32), col.astype(np.int32), name)


def values(nnz: int, n: int, kind: str = "int", dtype=np.float64, seed: int = 42):
    """(val, x) for one of the two value passes."""
    rng = np.random.default_rng(seed)
    if kind == "int":
        return (rng.integers(0, 10, size=nnz).astype(dtype), rng.integers(0, 10, size=n).astype(dtype))
    if kind == "real":
        return ((1.0 - rng.random(nnz)).astype(dtype), (1.0 - rng.random(n)).astype(dtype))
    raise ValueError(kind)


# ------------------------------------------------------------------------------------------------
# device generators (torch): full-size benchmark inputs built in HBM
# ------------------------------------------------------------------------------------------------

def device_values(nnz: int, n: int, kind: str, dtype, device, seed: int = 42):
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    if kind == "int":
        val = torch.randint(0, 10, (nnz,), generator=g, device=device, dtype=torch.int32).to(dtype)
        x = torch.randint(0, 10, (n,), generator=g, device=device, dtype=torch.int32).to(dtype)
    elif kind == "real":
        val = 1.0 - torch.rand(nnz, generator=g, device=device, dtype=dtype)
        x = 1.0 - torch.rand(n, generator=g, device=device, dtype=dtype)
    else:
        raise ValueError(kind)
    return val, x


def device_banded(m: int, per_row: int = 16, device="cuda", row_begin: int = 0, rows: int | None = None,
                  n: int | None = None):
    """Rows [row_begin, row_begin + rows) of the banded matrix with n columns (defaults: all of it).
    Returns (row_ptr, col) int32 tensors; row_ptr is rebased to 0."""
    import torch
    n = m if n is None else n
    rows = m - row_begin if rows is None else rows
    half = per_row // 2
    row_ptr = (torch.arange(rows + 1, device=device, dtype=torch.int64) * per_row).to(torch.int32)
    r = torch.arange(row_begin, row_begin + rows, device=device, dtype=torch.int64)
    col = (r[:, None] + torch.arange(-half, per_row - half, device=device, dtype=torch.int64)[None, :]) % n
    return row_ptr, col.reshape(-1).to(torch.int32)


def device_rmat(scale: int, edge_factor: int = 16, seed: int | None = None, device="cuda",
                abc=(0.57, 0.19, 0.19), chunk: int = 1 << 26):
    """R-MAT as ``rmat`` above, generated on the device in edge chunks (scale 25 = 537 M edges).
    Returns (row_ptr int32 (n+1), col int32 (nnz))."""
    import torch
    a, b, c = abc
    n = 1 << scale
    ne = edge_factor * n
    g = torch.Generator(device=device)
    g.manual_seed(scale if seed is None else seed)
    keys = []
    for s in range(0, ne, chunk):
        k = min(chunk, ne - s)
        key = torch.zeros(k, device=device, dtype=torch.int64)
        for bit in range(scale):
            u = torch.rand(k, generator=g, device=device, dtype=torch.float32)
            rbit = (u >= a + b).to(torch.int64)
            cbit = (((u >= a) & (u < a + b)) | (u >= a + b + c)).to(torch.int64)
            key |= (rbit << (bit + scale)) | (cbit << bit)
        keys.append(torch.unique(key))
        del key
    key = torch.unique(torch.cat(keys)) if len(keys) > 1 else keys[0]
    del keys
    row = key >> scale
    col = (key & (n - 1)).to(torch.int32)
    del key
    counts = torch.bincount(row, minlength=n)
    del row
    row_ptr = torch.zeros(n + 1, device=device, dtype=torch.int64)
    torch.cumsum(counts, 0, out=row_ptr[1:])
    return row_ptr.to(torch.int32), col


def device_laplacian27(nx: int, device="cuda", dtype=None, slab: int = 16):
    """27-point Laplacian on nx^3 (C4: nx = 320), built slab by slab along z to bound temporaries.
    Returns (row_ptr int32, col int32, val dtype)."""
    import torch
    dtype = torch.float32 if dtype is None else dtype
    m = nx ** 3
    nnz = (3 * nx - 2) ** 3
    col = torch.empty(nnz, device=device, dtype=torch.int32)
    val = torch.empty(nnz, device=device, dtype=dtype)
    row_cnt = torch.empty(m, device=device, dtype=torch.int64)
    d = torch.tensor([-1, 0, 1], device=device, dtype=torch.int64)
    dz, dy, dx = torch.meshgrid(d, d, d, indexing="ij")
    dz, dy, dx = dz.reshape(-1), dy.reshape(-1), dx.reshape(-1)
    doff = dx + dy * nx + dz * nx * nx
    w = 0
    for z0 in range(0, nx, slab):
        z1 = min(nx, z0 + slab)
        idx = torch.arange(z0 * nx * nx, z1 * nx * nx, device=device, dtype=torch.int64)
        ix, iy, iz = idx % nx, (idx // nx) % nx, idx // (nx * nx)
        ok = ((ix[:, None] + dx >= 0) & (ix[:, None] + dx < nx) & (iy[:, None] + dy >= 0) &
              (iy[:, None] + dy < nx) & (iz[:, None] + dz >= 0) & (iz[:, None] + dz < nx))
        cc = (idx[:, None] + doff)[ok]
        vv = torch.where((doff == 0)[None, :].expand_as(ok)[ok], 26.0, -1.0).to(dtype)
        k = cc.numel()
        col[w:w + k] = cc.to(torch.int32)
        val[w:w + k] = vv
        row_cnt[idx[0]:idx[-1] + 1] = ok.sum(1)
        w += k
        del idx, ix, iy, iz, ok, cc, vv
    assert w == nnz, (w, nnz)
    row_ptr = torch.zeros(m + 1, device=device, dtype=torch.int64)
    torch.cumsum(row_cnt, 0, out=row_ptr[1:])
    return row_ptr.to(torch.int32), col, val
