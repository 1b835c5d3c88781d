"""Row-range sharded CSR5 SpMV over the GPUs of one box (one process per GPU, torch.distributed).

The reference is single-device (no collective anywhere); this is the multi-GPU form
BASELINE.json's north_star asks for: the matrix is split into contiguous row ranges with balanced
nnz -- the same "row that holds nnz index b" search as the reference's tile partitioning
(generate_partition_pointer_s1_kernel, CSR5_cuda/detail/cuda/format_cuda.h:21-42) applied to the
boundaries b = g * nnz / G -- every rank builds its OWN CSR5 arrays for its rows through the
ordinary handle, x is replicated, and the y segments are concatenated on every rank.

Two exchange modes:

* ``"fused"`` (default): the concatenated y lives in symmetric memory (every rank's buffer mapped
  into every process over NVLink/NVSwitch) and ``csr5b200_spmv_scatter`` delivers this rank's rows to
  ALL of them -- through ONE store to the NVSwitch multicast address when the fabric offers it
  (``multicast=True``, the default when available), else through one store per peer.  ``scheme`` 1:
  the SpMV kernels themselves store each finished row to the destinations as tiles complete, so the
  all-gather traffic overlaps the tile stream; ``scheme`` 2: the SpMV runs on local memory and one
  coalesced pass pushes the segment (16-byte stores); 0 = auto by row structure.  One device-side
  barrier ends the step.  No NCCL call on the data path.
* ``"nccl"``: local SpMV into this rank's slot, then an all-gather(-v) over NCCL (the baseline the
  fused mode is measured against; also what runs on gloo in the CPU tests of the host logic).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

MAX_SCATTER = 8  # CSR5B200_MAX_SCATTER


# ---------------------------------------------------------------------------------------------
# host logic (device-agnostic; exercised on CPU with gloo in tests/test_sharded_cpu.py)
# ---------------------------------------------------------------------------------------------
def row_partition(row_ptr, parts: int) -> np.ndarray:
    """Boundaries r_0 = 0 <= r_1 <= ... <= r_G = m of G contiguous row ranges with balanced nnz:
    r_g = (number of rows r in [0, m] with row_ptr[r] <= g * nnz / G) - 1, i.e. the row that holds nnz
    index g*nnz/G, the LAST such row on ties (the rule of format_cuda.h:31-41 / utils_cuda.h:25-53).
    ``row_ptr`` may be a numpy array or a torch tensor (any device)."""
    if parts < 1:
        raise ValueError("parts must be >= 1")
    try:
        import torch
        is_t = isinstance(row_ptr, torch.Tensor)
    except ImportError:  # pragma: no cover
        is_t = False
    m = int(row_ptr.shape[0]) - 1
    nnz = int(row_ptr[-1])
    targets = [(g * nnz) // parts for g in range(parts + 1)]
    if is_t:
        import torch
        t = torch.tensor(targets, device=row_ptr.device, dtype=row_ptr.dtype)
        b = (torch.searchsorted(row_ptr.contiguous(), t, right=True) - 1).cpu().numpy().astype(np.int64)
    else:
        b = np.searchsorted(np.asarray(row_ptr), np.asarray(targets, dtype=np.asarray(row_ptr).dtype),
                            side="right").astype(np.int64) - 1
    b[0], b[-1] = 0, m
    return np.maximum.accumulate(np.clip(b, 0, m))


def shard_csr(row_ptr, col, val, row_begin: int, row_end: int):
    """Rows [row_begin, row_end) as a CSR of their own: (row_ptr rebased to 0, col slice, val slice).
    Slices are views; row_ptr is a new array/tensor of the same kind and dtype."""
    a, b = int(row_ptr[row_begin]), int(row_ptr[row_end])
    rp = row_ptr[row_begin:row_end + 1] - row_ptr[row_begin]
    return rp, col[a:b], val[a:b]


def allgather_v(y_full, bounds, rank: int, group=None):
    """Concatenate the ranks' y segments: on entry y_full[bounds[rank]:bounds[rank+1]] holds this rank's
    rows; on exit every rank holds all of y.  Equal segments use one all_gather_into_tensor (in place);
    ragged ones are padded to the longest segment (works on NCCL and gloo alike)."""
    import torch
    import torch.distributed as dist
    sizes = [int(bounds[g + 1] - bounds[g]) for g in range(len(bounds) - 1)]
    seg = y_full[int(bounds[rank]):int(bounds[rank + 1])]
    if len(set(sizes)) == 1:
        dist.all_gather_into_tensor(y_full, seg, group=group)
        return y_full
    mx = max(sizes)
    pad = torch.zeros(mx, dtype=y_full.dtype, device=y_full.device)
    pad[:seg.numel()] = seg
    tmp = torch.empty(len(sizes) * mx, dtype=y_full.dtype, device=y_full.device)
    dist.all_gather_into_tensor(tmp, pad, group=group)
    for g, n in enumerate(sizes):
        if g != rank and n:
            y_full[int(bounds[g]):int(bounds[g + 1])] = tmp[g * mx:g * mx + n]
    return y_full


# ---------------------------------------------------------------------------------------------
# the sharded handle (CUDA)
# ---------------------------------------------------------------------------------------------
class ShardedCsr5:
    """This rank's row range of a sharded matrix.  ``local_row_ptr`` is rebased to 0; ``bounds`` are
    the G + 1 global row boundaries (``row_partition``); ``n`` is the global column count."""

    def __init__(self, bounds, n: int, local_row_ptr, col, val, group=None, mode: str = "fused",
                 sigma: int = -1, multicast: bool | None = None, scheme: int = 0):
        import torch
        import torch.distributed as dist
        from . import handle as H
        self._torch, self._dist, self._H = torch, dist, H
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.bounds = np.asarray(bounds, np.int64)
        if len(self.bounds) != self.world + 1:
            raise ValueError("bounds must have world_size + 1 entries")
        if self.world > MAX_SCATTER and mode == "fused":
            raise ValueError(f"fused mode supports up to {MAX_SCATTER} ranks")
        self.m_global = int(self.bounds[-1])
        self.row_begin, self.row_end = int(self.bounds[self.rank]), int(self.bounds[self.rank + 1])
        self.m_local = self.row_end - self.row_begin
        self.n = int(n)
        self.dtype = val.dtype
        self.mode = mode
        self.h = H.anonymouslibHandle(self.m_local, self.n, self.dtype)
        err = self.h.inputCSR(int(col.numel()), local_row_ptr, col, val)
        if err:
            raise RuntimeError(self.h.error_string(err))
        self.h.setSigma(sigma)
        self.scheme = int(scheme)   # 0 auto, 1 stores fused into the SpMV kernels, 2 coalesced push pass
        self._resolved = False
        self._symm = None
        self._dst = None
        self.multicast = False
        dev = val.device
        if mode == "fused" and self.world > 1:
            import torch.distributed._symmetric_memory as symm_mem
            self.y_full = symm_mem.empty(self.m_global, dtype=self.dtype, device=dev)
            self._symm = symm_mem.rendezvous(self.y_full, group if group is not None else dist.group.WORLD)
            item = self.y_full.element_size()
            mc = int(self._symm.multicast_ptr or 0)
            ptrs = [int(p) + self.row_begin * item for p in self._symm.buffer_ptrs]
            self._dst_unicast = (C.c_void_p * self.world)(*ptrs)
            self._dst_multicast = (C.c_void_p * 1)(mc + self.row_begin * item) if mc else None
            self._want_multicast = multicast   # None = decide from the measured rule at the first spmv()
            self._dst = self._dst_unicast
        else:
            self.y_full = torch.empty(self.m_global, dtype=self.dtype, device=dev)
        self.y_local = self.y_full[self.row_begin:self.row_end]

    def setX(self, x) -> int:
        return self.h.setX(x)

    def asCSR5(self) -> int:
        return self.h.asCSR5()

    def _resolve_exchange(self):
        """Scheme and store kind of the fused mode, fixed at the first spmv() (the matrix must be in CSR5
        format).  Measured on 2 and 8 B200: rows stored by the SpMV kernels themselves win
        when every tile stores a run of consecutive rows (no empty rows, short rows) and unicast peer stores
        beat the multicast address for them; scattered row stores (dirty tiles, long rows) are better sent
        by the coalesced push pass, through the multicast address once 4 or more GPUs take part."""
        i = self.h.info()
        if self.scheme == 0:
            consecutive = not i.needs_zero_fill and (i.nnz // max(i.m, 1)) <= 64
            self.scheme = 1 if consecutive else 2
        self.h.set_option(self._H.OPT_EXCHANGE, self.scheme)
        want = self._want_multicast
        if want is None:
            want = self.scheme == 2 and self.world >= 4
        self.multicast = bool(want) and self._dst_multicast is not None
        self._dst = self._dst_multicast if self.multicast else self._dst_unicast
        self._resolved = True

    def spmv_local(self, alpha: float = 1.0) -> int:
        """Only this rank's rows (no exchange)."""
        return self.h.spmv(alpha, self.y_local)

    def spmv(self, alpha: float = 1.0):
        """y = alpha * A x, concatenated on every rank.  Returns the full y tensor (valid on the
        current stream once the call's work has completed)."""
        if self.world == 1:
            err = self.h.spmv(alpha, self.y_local)
        elif self._dst is not None:
            if not self._resolved:
                self._resolve_exchange()
            err = self.h.spmv_scatter(alpha, self.y_local, self._dst, len(self._dst), self.multicast)
            if not err:
                self._symm.barrier(channel=0)  # all peers' stores have landed before anyone reads y
        else:
            err = self.h.spmv(alpha, self.y_local)
            if not err:
                allgather_v(self.y_full, self.bounds, self.rank, self.group)
        if err:
            raise RuntimeError(self.h.error_string(err))
        return self.y_full

    def destroy(self) -> int:
        return self.h.destroy()

    def free(self):
        self.h.free()
