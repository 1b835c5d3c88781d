"""Host-side mirror of the reference's ``anonymouslibHandle<int, unsigned int, VT>``
(CSR5_cuda/anonymouslib_cuda.h:11-24) on top of the C ABI of ``libcsr5_b200.so``.

Same method names, argument meaning, call order and return codes as the reference class; arrays
are torch CUDA tensors (torch is only the owner of device memory and streams here).  Every compute
call goes through the C ABI -- there is no CPU or torch fallback: without the CUDA library the
constructor raises ``Csr5LibraryMissing``.

Reference call site this mirrors (CSR5_cuda/main.cu:59-108)::

    anonymouslibHandle<int, unsigned int, VALUE_TYPE> A(m, n);
    A.inputCSR(nnz, d_csrRowPtr, d_csrColIdx, d_csrVal);
    A.setX(d_x);
    A.setSigma(ANONYMOUSLIB_AUTO_TUNED_SIGMA);
    A.warmup();
    A.asCSR5();
    A.spmv(alpha, d_y);
    A.destroy();
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib

# detail/common.h:13-22, detail/cuda/common_cuda.h:11,15
ANONYMOUSLIB_SUCCESS = 0
ANONYMOUSLIB_UNKOWN_FORMAT = -1
ANONYMOUSLIB_UNSUPPORTED_CSR5_OMEGA = -2
ANONYMOUSLIB_CSR_TO_CSR5_FAILED = -3
ANONYMOUSLIB_UNSUPPORTED_CSR_SPMV = -4
ANONYMOUSLIB_UNSUPPORTED_VALUE_TYPE = -5
ANONYMOUSLIB_FORMAT_CSR = 0
ANONYMOUSLIB_FORMAT_CSR5 = 1
ANONYMOUSLIB_CSR5_OMEGA = 32
ANONYMOUSLIB_AUTO_TUNED_SIGMA = -1

OPT_KERNEL = 1
OPT_IGNORE_ALPHA = 2
OPT_TMA_STAGES = 3
OPT_TMA_WARPS = 4
OPT_CTAS_PER_SM = 5
OPT_KERNEL_TIMING = 6
OPT_DIRECT_WPB = 7
OPT_DIRECT_NCH = 8
OPT_HOT_COLUMNS = 9
OPT_HOT_THREADS = 10
OPT_EXCHANGE = 11
OPT_SIGMA_RULE = 12
OPT_DETERMINISTIC = 13
OPT_EXCHANGE_TRACE = 14
SIGMA_RULE_REFERENCE, SIGMA_RULE_B200 = 0, 1
EXCHANGE_AUTO, EXCHANGE_FUSED, EXCHANGE_PUSH = 0, 1, 2
# csr5b200_spmv_allgather transports (CSR5B200_TRANSPORT_*)
TRANSPORT_AUTO, TRANSPORT_COPY_ENGINE, TRANSPORT_SM_PUSH, TRANSPORT_SM_MULTICAST, TRANSPORT_IN_KERNEL, TRANSPORT_NONE = range(6)
TRANSPORT_NAMES = {"auto": 0, "ce": 1, "push": 2, "multicast": 3, "inkernel": 4, "none": 5}
EXCHANGE_TIMEOUT = -102
PROBES = {"stream": 1, "gather_nc": 2, "gather_cg": 3, "gather_ca": 4, "gather_only": 5, "fma_noseg": 6}
KERNEL_AUTO, KERNEL_DIRECT, KERNEL_TMA, KERNEL_HOT, KERNEL_TMA_PREFETCH = 0, 1, 2, 3, 4


def _ptr(t):
    return C.c_void_p(t.data_ptr() if t is not None and t.numel() > 0 else 0) if t is not None else C.c_void_p(0)


def _check_dev(t, dtype, name, count=None):
    import torch
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise TypeError(f"{name}: expected a CUDA tensor (device pointer, as in the reference)")
    if t.dtype != dtype:
        raise TypeError(f"{name}: expected dtype {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise ValueError(f"{name}: must be contiguous")
    if count is not None and t.numel() < count:
        raise ValueError(f"{name}: needs at least {count} elements, has {t.numel()}")


class anonymouslibHandle:
    """``anonymouslibHandle<int, unsigned int, VT>`` with VT chosen by ``dtype``
    (torch.float64 / torch.float32; VALUE_TYPE of CSR5_cuda/Makefile:4)."""

    def __init__(self, m: int, n: int, dtype=None):
        import torch
        self._torch = torch
        dtype = torch.float64 if dtype is None else dtype
        if dtype not in (torch.float64, torch.float32):
            raise TypeError("VALUE_TYPE must be float64 or float32")
        self._lib = _lib.load_library()
        self.dtype = dtype
        self.m, self.n = int(m), int(n)
        self._h = C.c_void_p()
        err = self._lib.csr5b200_create(self.m, self.n, 8 if dtype == torch.float64 else 4, C.byref(self._h))
        if err:
            raise RuntimeError(f"csr5b200_create: {self.error_string(err)}")
        self._keep = {}  # borrowed tensors are kept alive while the handle refers to them

    # -- the reference's public methods -----------------------------------------------------
    def warmup(self) -> int:
        return self._lib.csr5b200_warmup(self._h)

    def inputCSR(self, nnz: int, csr_row_pointer, csr_column_index, csr_value) -> int:
        t = self._torch
        _check_dev(csr_row_pointer, t.int32, "csr_row_pointer", self.m + 1)
        _check_dev(csr_column_index, t.int32, "csr_column_index", nnz)
        _check_dev(csr_value, self.dtype, "csr_value", nnz)
        self._keep.update(rp=csr_row_pointer, ci=csr_column_index, val=csr_value)
        return self._lib.csr5b200_input_csr(self._h, int(nnz), _ptr(csr_row_pointer),
                                            _ptr(csr_column_index), _ptr(csr_value))

    def asCSR(self) -> int:
        self._bind_stream()
        return self._lib.csr5b200_as_csr(self._h)

    def asCSR5(self) -> int:
        self._bind_stream()
        return self._lib.csr5b200_as_csr5(self._h)

    def setX(self, x) -> int:
        _check_dev(x, self.dtype, "x", self.n)
        self._keep["x"] = x
        return self._lib.csr5b200_set_x(self._h, _ptr(x))

    def spmv(self, alpha: float, y) -> int:
        _check_dev(y, self.dtype, "y", self.m)
        self._bind_stream()
        return self._lib.csr5b200_spmv(self._h, float(alpha), _ptr(y))

    def spmv_axpby(self, alpha: float, beta: float, y) -> int:
        """y = alpha * A * x + beta * y (csr5b200_spmv_axpby; the reference's commented-out `beta`,
        anonymouslib_cuda.h:281)."""
        _check_dev(y, self.dtype, "y", self.m)
        self._bind_stream()
        return self._lib.csr5b200_spmv_axpby(self._h, float(alpha), float(beta), _ptr(y))

    def spmv_allgather(self, alpha: float, beta: float, exchange: "_lib.Csr5Exchange") -> int:
        """One step of a row-range sharded SpMV with the y exchange overlapped (csr5b200_spmv_allgather)."""
        self._bind_stream()
        return self._lib.csr5b200_spmv_allgather(self._h, float(alpha), float(beta), C.byref(exchange))

    def exchange_status(self) -> int:
        return self._lib.csr5b200_exchange_status(self._h)

    def exchange_trace(self):
        """Timeline of the last traced step (OPT_EXCHANGE_TRACE): list of [tiles_done, carry_done, shipped] ms per row
        block in execution order, and the end of the step."""
        buf = (C.c_float * 256)()
        cnt = C.c_int(0)
        err = self._lib.csr5b200_exchange_trace(self._h, buf, 256, C.byref(cnt))
        if err:
            raise RuntimeError(self.error_string(err))
        v = [round(float(t), 4) for t in buf[:cnt.value]]
        return [v[i:i + 3] for i in range(0, len(v) - 1, 3)], (v[-1] if v else None)

    def spmv_scatter(self, alpha: float, y_local, y_dst, n_dst: int, multicast: bool = False) -> int:
        """Sharded mode (csr5b200_spmv_scatter): ``y_local`` is this shard's y segment (CUDA tensor in local
        memory), ``y_dst`` a ctypes array of ``n_dst`` device pointers, each the address of this shard's
        first row inside one destination's concatenated y -- or one multicast address."""
        _check_dev(y_local, self.dtype, "y_local", self.m)
        self._bind_stream()
        return self._lib.csr5b200_spmv_scatter(self._h, float(alpha), _ptr(y_local), int(n_dst),
                                               C.cast(y_dst, C.POINTER(C.c_void_p)), 1 if multicast else 0)

    def destroy(self) -> int:
        if not self._h:
            return ANONYMOUSLIB_SUCCESS
        self._bind_stream()
        return self._lib.csr5b200_destroy(self._h)

    def setSigma(self, sigma: int = ANONYMOUSLIB_AUTO_TUNED_SIGMA) -> None:
        self._lib.csr5b200_set_sigma(self._h, int(sigma))

    # -- additions --------------------------------------------------------------------------
    def set_option(self, option: int, value: int) -> int:
        return self._lib.csr5b200_set_option(self._h, option, value)

    def spmv_host(self, alpha: float, x_host, y_host) -> int:
        """x_host / y_host: numpy arrays or CPU tensors (pinned or pageable); H2D + spmv + D2H."""
        self._bind_stream()
        return self._lib.csr5b200_spmv_host(self._h, float(alpha), self._host_ptr(x_host, self.n),
                                            self._host_ptr(y_host, self.m))

    def spmv_host_batch(self, alpha: float, x_hosts, y_hosts) -> int:
        """Pipelined y_k = alpha * A * x_k for lists of host vectors (pinned for full overlap):
        upload of x_{k+1}, SpMV of x_k and download of y_{k-1} run concurrently."""
        if len(x_hosts) != len(y_hosts):
            raise ValueError("x_hosts and y_hosts must have the same length")
        self._bind_stream()
        k = len(x_hosts)
        xs = (C.c_void_p * k)(*[self._host_ptr(a, self.n) for a in x_hosts])
        ys = (C.c_void_p * k)(*[self._host_ptr(a, self.m) for a in y_hosts])
        return self._lib.csr5b200_spmv_host_batch(self._h, float(alpha), k, xs, ys)

    def probe(self, kind: int, repeats: int = 20) -> float:
        """csr5b200_probe: mean device ms of one stripped-down launch (PROBE_* kinds)."""
        ms = C.c_float(0)
        err = self._lib.csr5b200_probe(self._h, int(kind), int(repeats), C.byref(ms))
        if err:
            raise RuntimeError(self.error_string(err))
        return float(ms.value)

    def info(self) -> _lib.Csr5Info:
        out = _lib.Csr5Info()
        err = self._lib.csr5b200_get_info(self._h, C.byref(out))
        if err:
            raise RuntimeError(self.error_string(err))
        return out

    def meta_to_host(self) -> dict:
        """Copies of the CSR5 arrays (for the word-for-word comparison with the oracle)."""
        i = self.info()
        if i.format != ANONYMOUSLIB_FORMAT_CSR5:
            raise RuntimeError("handle is not in CSR5 format")
        n_tp = i.p + 1 if i.p else 0
        tile_ptr = np.zeros(n_tp, np.uint32)
        desc = np.zeros(i.p * 32 * i.num_packet, np.uint32)
        dop = np.zeros(n_tp, np.int32)
        doff = np.zeros(i.num_offsets, np.int32)
        cal = np.zeros(i.p, np.float64 if i.value_bytes == 8 else np.float32)
        err = self._lib.csr5b200_copy_meta_to_host(
            self._h, *(C.c_void_p(a.ctypes.data) if a.size else C.c_void_p(0)
                       for a in (tile_ptr, desc, dop, doff, cal)))
        if err:
            raise RuntimeError(self.error_string(err))
        return {"sigma": i.sigma, "bit_y": i.bit_y_offset, "bit_ss": i.bit_scansum_offset,
                "num_packet": i.num_packet, "p": i.p, "num_offsets": i.num_offsets,
                "tail_start": i.tail_partition_start, "tile_ptr": tile_ptr, "desc": desc,
                "desc_off_ptr": dop, "desc_off": doff, "calibrator": cal}

    def kernel_times_ms(self, capacity: int = 4096) -> np.ndarray:
        """Durations of the main SpMV kernel of the spmv() calls made with OPT_KERNEL_TIMING on."""
        buf = (C.c_float * capacity)()
        cnt = C.c_int(0)
        err = self._lib.csr5b200_get_kernel_times(self._h, buf, capacity, C.byref(cnt))
        if err:
            raise RuntimeError(self.error_string(err))
        return np.array(buf[:cnt.value], np.float64)

    def error_string(self, code: int) -> str:
        return self._lib.csr5b200_error_string(code).decode()

    def free(self) -> int:
        if not self._h:
            return ANONYMOUSLIB_SUCCESS
        self._bind_stream()
        err = self._lib.csr5b200_free(self._h)
        self._h = C.c_void_p()
        self._keep.clear()
        return err

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass

    # -- internals --------------------------------------------------------------------------
    def _bind_stream(self):
        s = self._torch.cuda.current_stream().cuda_stream
        self._lib.csr5b200_set_stream(self._h, C.c_void_p(s))

    def _host_ptr(self, a, count):
        t = self._torch
        if isinstance(a, t.Tensor):
            if a.is_cuda or a.dtype != self.dtype or not a.is_contiguous() or a.numel() < count:
                raise TypeError("host buffer: expected a contiguous CPU tensor of the handle's dtype")
            return C.c_void_p(a.data_ptr())
        want = np.float64 if self.dtype == t.float64 else np.float32
        if not isinstance(a, np.ndarray) or a.dtype != want or not a.flags.c_contiguous or a.size < count:
            raise TypeError("host buffer: expected a C-contiguous numpy array of the handle's dtype")
        return C.c_void_p(a.ctypes.data)


def call_anonymouslib(m, n, nnz, row_ptr, col, val, x, alpha: float = 1.0,
                      sigma: int = ANONYMOUSLIB_AUTO_TUNED_SIGMA) -> np.ndarray:
    """Host-array equivalent of the reference's ``call_anonymouslib`` (CSR5_cuda/main.cu:17-117)
    without the timing loop: upload, convert, ONE spmv, download y, restore, free."""
    lib = _lib.load_library()
    val = np.ascontiguousarray(val)
    if val.dtype not in (np.float64, np.float32):
        raise TypeError("VALUE_TYPE must be float64 or float32")
    row_ptr = np.ascontiguousarray(row_ptr, np.int32)
    col = np.ascontiguousarray(col, np.int32)
    x = np.ascontiguousarray(x, val.dtype)
    y = np.empty(m, val.dtype)
    err = lib.csr5b200_call_anonymouslib(
        m, n, nnz, row_ptr.ctypes.data, col.ctypes.data, val.ctypes.data, x.ctypes.data, y.ctypes.data,
        float(alpha), val.dtype.itemsize, int(sigma))
    if err:
        raise RuntimeError(f"csr5b200_call_anonymouslib: {lib.csr5b200_error_string(err).decode()}")
    return y
