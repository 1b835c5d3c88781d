"""Row-range sharded CSR5 SpMV over the GPUs of one box (one process per GPU, torch.distributed).

The reference is single-device (no collective anywhere); this is the multi-GPU form
BASELINE.json's north_star asks for: the matrix is split into contiguous row ranges with balanced
nnz -- the same "row that holds nnz index b" search as the reference's tile partitioning
(generate_partition_pointer_s1_kernel, CSR5_cuda/detail/cuda/format_cuda.h:21-42) applied to the
boundaries b = g * nnz / G -- every rank builds its OWN CSR5 arrays for its rows through the
ordinary handle, x is replicated, and the y segments are concatenated on every rank.

This module is a binding: the step itself is ``csr5b200_spmv_allgather`` of the C ABI
(csrc/csr5_exchange.cu); torch only provides the symmetric (peer-mapped) memory and the process group.

Exchange modes:

* ``"overlap"`` (default): the concatenated y lives in symmetric memory, TWO buffers used alternately.
  The shard's tiles are cut into row blocks that stream through the SpMV kernel while the finished
  blocks travel to the peers -- by the copy engines (``transport="ce"``), by a small grid of pushing
  CTAs (``"push"``), or through the NVSwitch multicast address (``"multicast"``).  Device-side flag
  barriers end the step; no NCCL call and no host synchronisation on the data path.  Because the
  buffers alternate, the y returned by step k stays valid until step k + 2 is enqueued, and a rank may
  run ahead into step k + 1 while its peers still read y_k (the write-after-read hazard of a single
  buffer).
* ``"fused"``: the first-generation scheme (``csr5b200_spmv_scatter``): every finished row is stored
  to all GPUs from inside the SpMV kernel (``scheme`` 1) or pushed by one coalesced pass afterwards
  (``scheme`` 2); kept for A/B measurements.  A barrier BEFORE the stores (all peers are done with the
  previous y) and one after them bracket the step.
* ``"nccl"``: local SpMV into this rank's slot, then an all-gather(-v) over NCCL (the baseline; also
  what runs on gloo in the CPU tests of the host logic).

The overlap step runs about a dozen streams per GPU and ends in a device-side barrier kernel that spins until the
peers arrive; export ``CUDA_DEVICE_MAX_CONNECTIONS=32`` before CUDA is initialised (``bench.py`` and the tests do) so
that every stream has a hardware queue of its own and no kernel is ever queued behind a barrier it does not depend
on.  The barriers time out (``timeout_ms``, default 20 s; ``exchange_status()``) rather than hang.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

MAX_SCATTER = 8  # CSR5B200_MAX_SCATTER


# ---------------------------------------------------------------------------------------------
# host logic (device-agnostic; exercised on CPU with gloo in tests/test_sharded_cpu.py)
# ---------------------------------------------------------------------------------------------
def row_partition(row_ptr, parts: int, row_cost: float = 0.0) -> np.ndarray:
    """Boundaries r_0 = 0 <= r_1 <= ... <= r_G = m of G contiguous row ranges with balanced nnz:
    r_g = (number of rows r in [0, m] with row_ptr[r] <= g * nnz / G) - 1, i.e. the row that holds nnz
    index g*nnz/G, the LAST such row on ties (the rule of format_cuda.h:31-41 / utils_cuda.h:25-53).
    ``row_ptr`` may be a numpy array or a torch tensor (any device).

    ``row_cost`` > 0: a shard's SpMV costs its non-zeros, its share of the y exchange -- which runs WHILE the SpMV
    runs -- costs its rows, ``row_cost`` non-zeros' worth each.  The ranges then minimise
    max over shards of max(nnz, row_cost * rows): the nnz rule wherever rows are not the bottleneck (few GPUs), fewer
    rows for the shards that hold the short and empty rows of a power-law matrix otherwise (R-MAT 25 on 8 GPUs by nnz
    alone: the last shard owns 14.6 M of the 33.5 M rows, and one GPU's multicast stream is what the step waits for)."""
    if parts < 1:
        raise ValueError("parts must be >= 1")
    if row_cost > 0:
        return _row_partition_minimax(row_ptr, parts, float(row_cost))
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


def _row_partition_minimax(row_ptr, parts: int, row_cost: float) -> np.ndarray:
    """Smallest T for which a left-to-right sweep that closes a range as late as nnz <= T and row_cost * rows <= T
    allow covers all rows with `parts` ranges (binary search on T; the sweep is `parts` binary searches)."""
    try:
        import torch
        if isinstance(row_ptr, torch.Tensor):
            row_ptr = row_ptr.detach().cpu().numpy()
    except ImportError:  # pragma: no cover
        pass
    rp = np.asarray(row_ptr, np.int64)
    m = rp.shape[0] - 1
    nnz = int(rp[-1])

    def sweep(T):
        b, r = [0], 0
        max_rows = int(T // row_cost)
        for _ in range(parts):
            if r >= m:
                b.append(m)
                continue
            r1 = int(np.searchsorted(rp, rp[r] + T, side="right")) - 1   # last boundary with nnz <= T
            r1 = min(r1, r + max_rows, m)
            if r1 <= r:
                return None                                             # one row alone exceeds T
            b.append(r1)
            r = r1
        return b if r >= m else None

    lo, hi = 1, int(nnz + row_cost * m) + 1
    while lo < hi:
        mid = (lo + hi) // 2
        if sweep(mid) is not None:
            hi = mid
        else:
            lo = mid + 1
    b = sweep(lo)
    return np.asarray(b, np.int64)


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

    def __init__(self, bounds, n: int, local_row_ptr, col, val, group=None, mode: str = "overlap",
                 sigma: int = -1, multicast: bool | None = None, scheme: int = 0, transport: str | int = "auto",
                 chunks: int = 0, push_ctas: int = 0, timeout_ms: int = 0, sigma_rule: int = 0,
                 push_threads: int = 0):
        import torch
        import torch.distributed as dist
        from . import _lib
        from . import handle as H
        self._torch, self._dist, self._H, self._lib = torch, dist, H, _lib
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.bounds = np.asarray(bounds, np.int64)
        if len(self.bounds) != self.world + 1:
            raise ValueError("bounds must have world_size + 1 entries")
        if mode not in ("overlap", "fused", "nccl"):
            raise ValueError(f"unknown exchange mode {mode!r}")
        if self.world > MAX_SCATTER and mode != "nccl":
            raise ValueError(f"{mode} mode supports up to {MAX_SCATTER} ranks")
        self.m_global = int(self.bounds[-1])
        self.row_begin, self.row_end = int(self.bounds[self.rank]), int(self.bounds[self.rank + 1])
        self.m_local = self.row_end - self.row_begin
        self.n = int(n)
        self.dtype = val.dtype
        self.mode = mode if self.world > 1 else "local"
        self.h = H.anonymouslibHandle(self.m_local, self.n, self.dtype)
        err = self.h.inputCSR(int(col.numel()), local_row_ptr, col, val)
        if err:
            raise RuntimeError(self.h.error_string(err))
        self.h.set_option(H.OPT_SIGMA_RULE, int(sigma_rule))
        self.h.setSigma(sigma)
        self.scheme = int(scheme)   # fused mode: 0 auto, 1 stores fused into the SpMV kernels, 2 coalesced push pass
        self.transport = H.TRANSPORT_NAMES[transport] if isinstance(transport, str) else int(transport)
        self.chunks, self.push_ctas, self.timeout_ms = int(chunks), int(push_ctas), int(timeout_ms)
        self.push_threads = int(push_threads)
        self._resolved = False
        self._symm = None
        self._dst = None
        self._parity = 0
        self.multicast = False
        dev = val.device
        item = torch.empty(0, dtype=self.dtype).element_size()
        if self.mode == "overlap":
            import torch.distributed._symmetric_memory as symm_mem
            grp = group if group is not None else dist.group.WORLD
            self._stride = (self.m_global + 31) // 32 * 32          # second buffer starts 128-byte aligned
            self._ybuf = symm_mem.empty(2 * self._stride, dtype=self.dtype, device=dev)
            self._symm = symm_mem.rendezvous(self._ybuf, grp)
            self._flagbuf = symm_mem.empty(64, dtype=torch.int32, device=dev)
            self._flagbuf.zero_()
            self._fsymm = symm_mem.rendezvous(self._flagbuf, grp)
            torch.cuda.synchronize(dev)
            dist.barrier(group)                                     # every rank's flag words are zero before anyone signals
            mc = int(self._symm.multicast_ptr or 0)
            self._ex = []
            for b in range(2):
                ex = _lib.Csr5Exchange()
                ex.rank, ex.world = self.rank, self.world
                for k in range(self.world):
                    ex.y_full[k] = int(self._symm.buffer_ptrs[k]) + b * self._stride * item
                    ex.flags[k] = int(self._fsymm.buffer_ptrs[k])
                ex.y_multicast = (mc + b * self._stride * item) if mc else None
                ex.row_begin = self.row_begin
                ex.entry_barrier = 0                                # two alternating buffers: see the module docstring
                self._ex.append(ex)
            self.has_multicast = bool(mc)
            if self.transport == H.TRANSPORT_SM_MULTICAST and not mc:
                raise RuntimeError("transport 'multicast': the fabric offers no multicast address")
            self.y_full = self._ybuf[:self.m_global]
        elif self.mode == "fused":
            import torch.distributed._symmetric_memory as symm_mem
            self.y_full = symm_mem.empty(self.m_global, dtype=self.dtype, device=dev)
            self._symm = symm_mem.rendezvous(self.y_full, group if group is not None else dist.group.WORLD)
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

    def _views(self, parity: int):
        base = parity * self._stride
        full = self._ybuf[base:base + self.m_global]
        return full, full[self.row_begin:self.row_end]

    def spmv_local(self, alpha: float = 1.0) -> int:
        """Only this rank's rows (no exchange)."""
        return self.h.spmv(alpha, self.y_local)

    def spmv(self, alpha: float = 1.0, beta: float = 0.0):
        """y = alpha * A x (+ beta * y), concatenated on every rank.  Returns the full y tensor (valid on the
        current stream once the call's work has completed; in overlap mode until the next-but-one call)."""
        if self.mode == "overlap":
            ex = self._ex[self._parity]
            ex.transport, ex.chunks, ex.push_ctas, ex.timeout_ms = self.transport, self.chunks, self.push_ctas, self.timeout_ms
            ex.push_threads = self.push_threads
            if beta != 0.0 and self._parity != self._last_parity():
                # beta * y refers to the y of the previous step, which lives in the other buffer
                prev_full, prev_local = self._views(self._last_parity())
                self._views(self._parity)[1].copy_(prev_local)
            err = self.h.spmv_allgather(alpha, beta, ex)
            self.y_full, self.y_local = self._views(self._parity)
            self._prev = self._parity
            self._parity ^= 1
        elif self.mode == "fused":
            if beta != 0.0:
                raise ValueError("fused mode has no beta term; use mode='overlap'")
            if not self._resolved:
                self._resolve_exchange()
            self._symm.barrier(channel=1)      # every peer is done reading the previous y before it is overwritten
            err = self.h.spmv_scatter(alpha, self.y_local, self._dst, len(self._dst), self.multicast)
            if not err:
                self._symm.barrier(channel=0)  # all peers' stores have landed before anyone reads y
        elif self.mode == "nccl":
            err = self.h.spmv_axpby(alpha, beta, self.y_local) if beta != 0.0 else self.h.spmv(alpha, self.y_local)
            if not err:
                allgather_v(self.y_full, self.bounds, self.rank, self.group)
        else:
            err = self.h.spmv_axpby(alpha, beta, self.y_local) if beta != 0.0 else self.h.spmv(alpha, self.y_local)
        if err:
            raise RuntimeError(self.h.error_string(err))
        return self.y_full

    def _last_parity(self) -> int:
        return getattr(self, "_prev", self._parity)

    def iterate(self, steps: int, alpha: float = 1.0):
        """x_{k+1} = alpha * A x_k for `steps` steps without leaving the devices: the gathered y of one step is
        the x of the next (square matrices).  In overlap mode the two y buffers alternate as x and y, and the
        exchange of step k overlaps its own SpMV; returns the last y."""
        if self.m_global != self.n:
            raise ValueError("iterate() needs a square matrix (y is fed back as x)")
        y = None
        for _ in range(int(steps)):
            y = self.spmv(alpha)
            if self.mode in ("fused", "nccl", "local"):
                y = y.clone()   # one buffer: the next step overwrites it while it is being read as x
            err = self.h.setX(y)
            if err:
                raise RuntimeError(self.h.error_string(err))
        if y is not None and self.mode == "overlap":
            # leave x in a tensor of its own: the y buffer it lives in is rewritten by the step after next
            self._x_keep = y.clone()
            err = self.h.setX(self._x_keep)
            if err:
                raise RuntimeError(self.h.error_string(err))
        return y

    def exchange_status(self) -> int:
        """Synchronises; EXCHANGE_TIMEOUT if a device-side barrier gave up waiting for a peer."""
        return self.h.exchange_status() if self.mode == "overlap" else 0

    def destroy(self) -> int:
        return self.h.destroy()

    def free(self):
        self.h.free()


# ---------------------------------------------------------------------------------------------
# binding of the single-process C++ host API (include/csr5_b200_sharded.h)
# ---------------------------------------------------------------------------------------------
BARRIER_AUTO, BARRIER_FLAGS, BARRIER_EVENTS = 0, 1, 2


class ShardedCsr5Native:
    """``csr5b200_sharded_*``: all shards driven from this one process (one worker thread per shard inside the
    library).  ``devices[s]`` is the CUDA device of shard s; listing a device twice puts two shards on it (how the
    single-GPU test box exercises the path).  Host numpy arrays in, host numpy arrays out; no torch involved."""

    def __init__(self, devices, dtype=np.float64):
        from . import _lib
        self._lib = _lib.load_library()
        self.dtype = np.dtype(dtype)
        if self.dtype not in (np.dtype(np.float64), np.dtype(np.float32)):
            raise TypeError("VALUE_TYPE must be float64 or float32")
        self.devices = [int(d) for d in devices]
        self._s = C.c_void_p()
        arr = (C.c_int * len(self.devices))(*self.devices)
        self._check(self._lib.csr5b200_sharded_create(len(self.devices), arr, self.dtype.itemsize, C.byref(self._s)))
        self.m = self.n = 0

    def _check(self, err):
        if err:
            raise RuntimeError(f"csr5b200_sharded: {self._lib.csr5b200_error_string(err).decode()} ({err})")

    def inputCSR(self, m, n, row_ptr, col, val):
        rp = np.ascontiguousarray(row_ptr, np.int32)
        ci = np.ascontiguousarray(col, np.int32)
        v = np.ascontiguousarray(val, self.dtype)
        self.m, self.n = int(m), int(n)
        self._check(self._lib.csr5b200_sharded_input_csr_host(self._s, self.m, self.n, int(ci.size), rp.ctypes.data,
                                                              ci.ctypes.data, v.ctypes.data))

    def set_partition(self, row_cost=0.0):
        """Before inputCSR: shards minimise max(nnz, row_cost * rows) (row_partition's rule)."""
        self._check(self._lib.csr5b200_sharded_set_partition(self._s, float(row_cost)))

    def setSigma(self, sigma=-1):
        self._check(self._lib.csr5b200_sharded_set_sigma(self._s, int(sigma)))

    def set_option(self, option, value):
        self._check(self._lib.csr5b200_sharded_set_option(self._s, int(option), int(value)))

    def set_exchange(self, transport=0, chunks=0, push_ctas=0, barrier=BARRIER_AUTO, timeout_ms=0):
        from . import handle as H
        t = H.TRANSPORT_NAMES[transport] if isinstance(transport, str) else int(transport)
        self._check(self._lib.csr5b200_sharded_set_exchange(self._s, t, int(chunks), int(push_ctas), int(barrier),
                                                            int(timeout_ms)))

    def setX(self, x):
        xx = np.ascontiguousarray(x, self.dtype)
        if xx.size < self.n:
            raise ValueError("x: needs n values")
        self._check(self._lib.csr5b200_sharded_set_x_host(self._s, xx.ctypes.data))

    def asCSR5(self):
        self._check(self._lib.csr5b200_sharded_as_csr5(self._s))

    def spmv(self, alpha=1.0, beta=0.0):
        self._check(self._lib.csr5b200_sharded_spmv(self._s, float(alpha), float(beta)))

    def iterate(self, steps, alpha=1.0):
        self._check(self._lib.csr5b200_sharded_iterate(self._s, int(steps), float(alpha)))

    def synchronize(self):
        self._check(self._lib.csr5b200_sharded_synchronize(self._s))

    def y(self, shard=0) -> np.ndarray:
        """The concatenated y of the last step as shard `shard`'s device holds it (synchronises)."""
        out = np.empty(self.m, self.dtype)
        self._check(self._lib.csr5b200_sharded_copy_y_to_host(self._s, int(shard), out.ctypes.data))
        return out

    def bounds(self) -> np.ndarray:
        b = (C.c_longlong * (len(self.devices) + 1))()
        self._check(self._lib.csr5b200_sharded_get_bounds(self._s, b))
        return np.array(list(b), np.int64)

    def shard_info(self, shard):
        from . import _lib
        h = C.c_void_p()
        self._check(self._lib.csr5b200_sharded_get_handle(self._s, int(shard), C.byref(h)))
        out = _lib.Csr5Info()
        self._check(self._lib.csr5b200_get_info(h, C.byref(out)))
        return out

    def destroy(self):
        if self._s:
            self._lib.csr5b200_sharded_destroy(self._s)
            self._s = C.c_void_p()

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass
