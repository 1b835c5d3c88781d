#!/usr/bin/env python
"""bench.py -- CSR5 SpMV on B200: GFLOPS and achieved HBM GB/s against the streaming roofline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c3|c4|c5] [--impl ours|reference]

A step is one SpMV (y = A x) over the resident matrix, the unit the reference times in its
benchmark loop (CSR5_cuda/main.cu:93-106: NUM_RUN back-to-back spmv() between two events).

N = 1: workload = BASELINE.json configs[1] (banded 10M x 10M, 16 nnz/row, FP64); the other single-GPU
configurations (configs[2] R-MAT 22, configs[3] Laplacian 320^3 FP32) are measured in the same run and
reported under `extra_workloads` (value, kernel time, roofline fraction, parity).
N > 1: one process per GPU (torchrun), weak scaling -- rank g owns the row range [g*m, (g+1)*m) of the
N*m-row banded matrix (its own CSR5 arrays), x is replicated, and every step ends with every rank holding
all of y; the exchange is overlapped with the SpMV (csr5b200_spmv_allgather).  The same
line carries `multi_gpu.c5_strong`: BASELINE.json configs[4] (R-MAT 25 split by balanced nnz) on the N
GPUs against the same matrix on one.

One JSON line on stdout (rank 0).  `value` = whole-job GFLOPS with everything resident in HBM;
`e2e` = the same with every step's x uploaded from and y downloaded to pinned host memory
(csr5b200_spmv_host_batch on one GPU; a four-stream pipeline around ShardedCsr5.spmv on several);
`roofline` = algorithmic bytes / live CUDA-event duration of the main SpMV kernel against the measured
HBM copy bandwidth; `cpu_baseline` = the reference's own CSR5_avx2 backend (oracle/_ref, compiled from
/root/reference) on this box's host cores on a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# one hardware queue per stream: the overlapped exchange runs ~16 streams per GPU, and a device-side barrier must
# never sit in a queue in front of work it (transitively) waits for
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

WORKLOADS = {
    "c2": "banded 10M x 10M, 16 nnz/row, FP64 (BASELINE.json configs[1])",
    "c3": "R-MAT scale 22, edge factor 16, FP64 (BASELINE.json configs[2])",
    "c4": "27-pt Laplacian 320^3, FP32 (BASELINE.json configs[3])",
    "c5": "R-MAT scale 25, edge factor 16, FP64, row-range sharded by balanced nnz (BASELINE.json configs[4])",
}
ROW_COST = [8.0]   # --row-cost: strong-scaling workloads minimise max(nnz, cost * rows) per shard instead of balancing nnz alone
STRONG = {"c3", "c5"}   # fixed matrix split over the ranks (strong scaling); c2 grows with N (weak)


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# stdout carries exactly ONE JSON line: everything else that native libraries print to file descriptor 1
# (e.g. "NCCL version ...") is sent to stderr by pointing fd 1 at fd 2 for the duration of the run.
_REAL_STDOUT = None


def capture_stdout():
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(obj):
    line = (json.dumps(obj) + "\n").encode()
    sys.stdout.flush()
    if _REAL_STDOUT is not None:
        os.write(_REAL_STDOUT, line)
    else:
        os.write(1, line)


# ---------------------------------------------------------------------------------------------
# clocks: sampled DURING the timed regions
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    """NVML sampler thread (falls back to one-shot nvidia-smi queries)."""
    REASONS = {
        0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
        0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
        0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting",
    }

    def __init__(self, index: int, period_s: float = 0.004, uuid: str | None = None):
        """index: CUDA device ordinal of this process; uuid: that device's UUID.  NVML numbers the
        physical GPUs and ignores CUDA_VISIBLE_DEVICES, so the device is looked up by UUID first."""
        self.index, self.period = index, period_s
        self.samples = []  # (t, sm_mhz, reasons bitmask)
        self.max_mhz = None
        self._stop = threading.Event()
        self._thr = None
        self._nvml = None
        self._smi_id = str(index)
        cvd = os.environ.get("CUDA_VISIBLE_DEVICES", "")
        ids = [t.strip() for t in cvd.split(",") if t.strip()]
        if uuid:
            self._smi_id = uuid if uuid.startswith("GPU-") else "GPU-" + uuid
        elif index < len(ids):
            self._smi_id = ids[index]
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nvml = pynvml
            try:
                if not self._smi_id.startswith("GPU-"):
                    raise ValueError("no uuid")
                self._dev = pynvml.nvmlDeviceGetHandleByUUID(self._smi_id)
            except Exception:
                phys = int(self._smi_id) if self._smi_id.isdigit() else index
                self._dev = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self._dev, pynvml.NVML_CLOCK_SM))
        except Exception as e:  # pragma: no cover
            log(f"[bench] NVML unavailable ({e}); falling back to nvidia-smi")

    def _one(self):
        if self._nvml is not None:
            n = self._nvml
            mhz = float(n.nvmlDeviceGetClockInfo(self._dev, n.NVML_CLOCK_SM))
            try:
                rs = int(n.nvmlDeviceGetCurrentClocksEventReasons(self._dev))
            except Exception:
                rs = int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self._dev))
            return mhz, rs
        out = subprocess.run(
            ["nvidia-smi", "-i", self._smi_id, "--query-gpu=clocks.sm,clocks.max.sm,clocks_event_reasons.active",
             "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=10).stdout.strip()
        a = [t.strip() for t in out.split(",")]
        self.max_mhz = float(a[1])
        return float(a[0]), int(a[2], 16)

    def _run(self):
        while not self._stop.is_set():
            try:
                mhz, rs = self._one()
                self.samples.append((time.perf_counter(), mhz, rs))
            except Exception:
                pass
            self._stop.wait(self.period)

    def start(self):
        self._thr = threading.Thread(target=self._run, daemon=True)
        self._thr.start()

    def stop(self):
        self._stop.set()
        if self._thr:
            self._thr.join(timeout=5)

    def summary(self, windows):
        """windows: list of (t0, t1) host timestamps of the timed regions."""
        inside = [(m, r) for (t, m, r) in self.samples if any(a <= t <= b for a, b in windows)]
        used = inside if inside else [(m, r) for (_, m, r) in self.samples]
        mask = 0
        for _, r in used:
            mask |= r
        reasons = sorted(name for bit, name in self.REASONS.items() if mask & bit and name != "gpu_idle")
        return {"sm_mhz": float(np.median([m for m, _ in used])) if used else None,
                "sm_max_mhz": self.max_mhz, "reasons": reasons, "samples": len(used),
                "sampled": "in timed regions" if inside else "whole run"}


# ---------------------------------------------------------------------------------------------
# workloads
# ---------------------------------------------------------------------------------------------
def build_workload(name, torch, device, rank, world):
    """-> dict(row_ptr, col, val, x, m (local rows), n, row_begin, dtype, desc)"""
    from benchmark_spmv_using_csr5_b200 import matrices as M
    if name == "c2":
        m_local = 10_000_000
        n = m_local * world
        rp, ci = M.device_banded(n, 16, device, row_begin=rank * m_local, rows=m_local, n=n)
        dtype = torch.float64
        val, _ = M.device_values(ci.numel(), 1, "real", dtype, device, seed=42 + rank)
        _, x = M.device_values(1, n, "real", dtype, device, seed=4242)  # same x on every rank
        return dict(row_ptr=rp, col=ci, val=val, x=x, m=m_local, n=n, dtype=dtype,
                    bounds=np.arange(world + 1, dtype=np.int64) * m_local)
    if name in ("c3", "c5"):
        # every rank generates the same seeded matrix and keeps its nnz-balanced row range
        from benchmark_spmv_using_csr5_b200 import sharded as S
        rp, ci = M.device_rmat(22 if name == "c3" else 25, device=device)
        n = rp.numel() - 1
        dtype = torch.float64
        val, x = M.device_values(ci.numel(), n, "real", dtype, device, seed=42)
        bounds = S.row_partition(rp, world, row_cost=ROW_COST[0])
        if world > 1:
            lrp, lci, lval = S.shard_csr(rp, ci, val, int(bounds[rank]), int(bounds[rank + 1]))
            lrp, lci, lval = lrp.contiguous(), lci.clone(), lval.clone()
            del rp, ci, val
            torch.cuda.empty_cache()
            rp, ci, val = lrp, lci, lval
        return dict(row_ptr=rp, col=ci, val=val, x=x, m=int(bounds[rank + 1] - bounds[rank]), n=n, dtype=dtype,
                    bounds=bounds)
    if world != 1:
        raise SystemExit(f"workload {name} is a single-GPU configuration")
    if name == "c4":
        rp, ci, val = M.device_laplacian27(320, device=device, dtype=torch.float32)
        n = rp.numel() - 1
        dtype = torch.float32
        _, x = M.device_values(1, n, "real", dtype, device, seed=42)
        return dict(row_ptr=rp, col=ci, val=val, x=x, m=n, n=n, dtype=dtype, bounds=np.array([0, n], np.int64))
    raise SystemExit(f"unknown workload {name}")


def algorithmic_bytes(m, n, nnz, vb):
    """compulsory CSR traffic, each array touched once."""
    return nnz * (vb + 4) + (m + 1) * 4 + n * vb + m * vb


def reference_getB(m, nnz, vb):
    """detail/utils.h:10-14 (counts one x read per nnz) -- for comparability with the reference's print."""
    return (m + 1 + nnz) * 4 + (2 * nnz + m) * vb


def reference_rows(torch, w):
    """Per-row FP64 sums of this rank's rows evaluated on the device by an independent route (torch index_add_ of
    the FP64 products onto their row; col/val must be in CSR order), and sum_j |a_ij x_j| -- the scale of the usual
    row-wise backward-error normalisation (for the all-positive C2/C3 inputs: the plain relative error)."""
    m = w["row_ptr"].numel() - 1
    counts = (w["row_ptr"][1:] - w["row_ptr"][:-1]).long()
    rows = torch.repeat_interleave(torch.arange(m, device=counts.device), counts)
    prod = w["val"].double() * w["x"].double()[w["col"].long()]
    ref = torch.zeros(m, device=prod.device, dtype=torch.float64).index_add_(0, rows, prod)
    scale = torch.zeros(m, device=prod.device, dtype=torch.float64).index_add_(0, rows, prod.abs_())
    return ref, scale


def host_threads():
    """Host cores this process may use (the container's cpuset), NOT the OpenMP default: torchrun exports
    OMP_NUM_THREADS=1, which must not neuter the CPU arm."""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:  # pragma: no cover
        return os.cpu_count() or 1


# ---------------------------------------------------------------------------------------------
# CPU baseline: the reference's CSR5_avx2 backend on the host cores (bounded sample)
# ---------------------------------------------------------------------------------------------
def cpu_reference_run(workload, warmup, runs, sample_rows=None):
    """Times oracle/_ref/libref_avx2.so (the reference's own CSR5_avx2, all host threads) on a bounded
    sample of `workload`.  Returns dict(value GFLOPS, ms, cores, kind, sample, nnz)."""
    import oracle
    from benchmark_spmv_using_csr5_b200 import matrices as M
    nthreads = host_threads()
    if workload == "c2":
        rows = sample_rows or 2_500_000
        A = M.banded(rows, 16)
        val, x = M.values(A.nnz, A.n, "real", np.float64, 42)
        sample = f"banded {rows} x {rows}, 16 nnz/row, FP64 ({A.nnz} nnz = 1/{10_000_000 // rows} of the workload's rows)"
    elif workload in ("c3", "c5"):
        A = M.rmat(18)
        val, x = M.values(A.nnz, A.n, "real", np.float64, 42)
        sample = f"R-MAT scale 18 ({A.nnz} nnz), FP64"
    elif workload == "c4":
        A, val = M.laplacian27(128)
        val = val.astype(np.float32)
        _, x = M.values(1, A.n, "real", np.float32, 42)
        ms, _y, threads = oracle.ref_csr_omp_bench_f32(A.m, A.row_ptr, A.col, val, x, nthreads, warmup, runs)
        return dict(value=2.0 * A.nnz / (ms * 1e6), ms=ms, cores=threads, kind="port", nnz=A.nnz,
                    sample=f"27-pt Laplacian 128^3 FP32 ({A.nnz} nnz), OpenMP scalar CSR loop "
                           "(the reference has no FP32 AVX2 path, README.md:36)")
    else:
        raise SystemExit(workload)
    if not oracle.ref_available():
        raise SystemExit("oracle/_ref/libref_avx2.so missing: run __graft_entry__.build() where /root/reference exists")
    ms, conv_ms, y, threads = oracle.ref_avx2_bench(A.m, A.n, A.row_ptr, A.col, val, x, nthreads, warmup, runs)
    y_ref = oracle.csr_spmv(A.m, A.row_ptr, A.col, val, x)
    ok = np.allclose(y, y_ref, rtol=1e-10, atol=0)
    return dict(value=2.0 * A.nnz / (ms * 1e6), ms=ms, cores=threads, kind="reference", nnz=A.nnz,
                sample=sample + f"; CSR5_avx2 sigma 16 omega 4, {threads} OpenMP threads; {warmup} warm-up + {runs} timed SpMVs"
                                f"; y check vs scalar CSR: {'pass' if ok else 'FAIL'}; CSR->CSR5 {conv_ms:.1f} ms")


def run_reference_arm(args):
    """--impl reference: the reference's own CPU implementation of the path on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t0 = time.time()
    r = cpu_reference_run(args.workload, max(args.warmup, 1), max(args.steps, 1))
    out = {
        "impl": "reference",
        "metric": "FP64 SpMV GFLOPS" if args.workload != "c4" else "FP32 SpMV GFLOPS",
        "value": r["value"], "unit": "GFLOP/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": r["ms"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64" if args.workload != "c4" else "f32", "data": "synthetic",
        "config": {"workload": WORKLOADS[args.workload], "sample": r["sample"]},
        "cpu_baseline": {"value": r["value"], "unit": "GFLOP/s", "cores": r["cores"], "kind": r["kind"],
                         "sample": r["sample"]},
        "e2e": {"value": r["value"], "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": time.time() - t0,
    }
    emit(out)


# ---------------------------------------------------------------------------------------------
# NUMA: bind a rank (threads + the pinned buffers it allocates afterwards) to its GPU's node
# ---------------------------------------------------------------------------------------------
def bind_to_gpu_numa(torch, local_rank):
    """Best effort.  Returns a dict describing what happened (goes into the JSON line)."""
    info = {"bound": False}
    try:
        bus = torch.cuda.get_device_properties(local_rank).pci_bus_id
        dom = getattr(torch.cuda.get_device_properties(local_rank), "pci_domain_id", 0)
        dev = getattr(torch.cuda.get_device_properties(local_rank), "pci_device_id", 0)
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev:02x}.0/numa_node"
        node = int(open(path).read().strip())
        info["gpu_numa_node"] = node
        if node < 0:
            return info
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        allowed = os.sched_getaffinity(0)
        use = sorted(cpus & allowed)
        info["node_cpus"], info["allowed_cpus"] = len(cpus), len(allowed)
        if use:
            os.sched_setaffinity(0, use)
            info["bound"] = True
            info["cpus_used"] = len(use)
    except Exception as e:  # pragma: no cover
        info["error"] = f"{type(e).__name__}: {e}"
    return info


# ---------------------------------------------------------------------------------------------
# one workload on one GPU
# ---------------------------------------------------------------------------------------------
KERNEL_NAMES = {1: "spmv_direct_kernel", 2: "spmv_tma_kernel", 3: "spmv_hot_kernel", 4: "spmv_tma_kernel<prefetch>"}


def hbm_peak():
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        return float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy)"
    return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"


def apply_tuning(A, H, args):
    A.set_option(H.OPT_KERNEL, args.kernel)
    A.set_option(H.OPT_TMA_STAGES, args.stages)
    A.set_option(H.OPT_TMA_WARPS, args.warps)
    A.set_option(H.OPT_CTAS_PER_SM, args.ctas_per_sm)
    A.set_option(H.OPT_HOT_COLUMNS, args.hot)
    A.set_option(H.OPT_HOT_THREADS, args.hot_threads)
    A.set_option(H.OPT_DIRECT_WPB, args.wpb)
    A.set_option(H.OPT_DIRECT_NCH, args.nch)
    A.set_option(H.OPT_SIGMA_RULE, args.sigma_rule)


def timed_loop(torch, fn, warmup, steps, windows=None, sync=None):
    """W untimed + K timed calls of fn between two events on the current stream, device-synchronised (and, through
    `sync`, barriered across ranks) on both sides.  Returns ms per step on this rank."""
    sync = sync or torch.cuda.synchronize
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(warmup):
        fn()
    sync()
    h0 = time.perf_counter()
    ev0.record()
    for _ in range(steps):
        fn()
    ev1.record()
    sync()
    if windows is not None:
        windows.append((h0, time.perf_counter()))
    return ev0.elapsed_time(ev1) / steps


def measure_single(torch, H, name, args, device, windows, steps, full):
    """Converts, checks and times workload `name` on one GPU.  full = also the end-to-end (host buffer) paths."""
    w = build_workload(name, torch, device, 0, 1)
    m, n, nnz, dtype = w["m"], w["n"], w["col"].numel(), w["dtype"]
    vb = 8 if dtype == torch.float64 else 4
    A = H.anonymouslibHandle(m, n, dtype)
    assert A.inputCSR(nnz, w["row_ptr"], w["col"], w["val"]) == 0
    assert A.setX(w["x"]) == 0
    apply_tuning(A, H, args)   # before setSigma: the rule behind AUTO is one of the options
    A.setSigma(args.sigma)
    A.warmup()
    torch.cuda.synchronize()
    conv = []
    for it in range(6):   # first call pays the module loads; then the median of 5
        t0 = time.perf_counter()
        err = A.asCSR5()
        torch.cuda.synchronize()
        conv.append((time.perf_counter() - t0) * 1e3)
        assert err == 0, A.error_string(err)
        if it < 5:
            assert A.asCSR() == 0
    info = A.info()
    y = torch.full((m,), float("nan"), device=device, dtype=dtype)

    # one checked SpMV (the reference checks its first call, main.cu:80-82, 360-384)
    assert A.spmv(1.0, y) == 0
    torch.cuda.synchronize()
    A.asCSR()   # the check needs col/val in CSR order
    ref, scale = reference_rows(torch, w)
    max_rel = float(((y.double() - ref).abs() / scale.clamp_min(1e-300)).max())
    del ref, scale
    assert A.asCSR5() == 0
    tol = 1e-6 if vb == 8 else 1e-4
    assert max_rel <= tol, f"{name}: parity check failed (max rel err {max_rel})"

    for _ in range(max(args.warmup, 3)):
        A.spmv(1.0, y)
    A.kernel_times_ms()  # drop
    A.set_option(H.OPT_KERNEL_TIMING, 1)
    ms_step = timed_loop(torch, lambda: A.spmv(1.0, y), 0, steps, windows)
    kt = A.kernel_times_ms()
    A.set_option(H.OPT_KERNEL_TIMING, 0)
    info = A.info()
    peak, peak_src = hbm_peak()
    b_alg = algorithmic_bytes(m, n, nnz, vb)
    k_ms = float(kt.mean()) if kt.size else ms_step
    achieved = b_alg / (k_ms * 1e6)
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": None, "kernel": KERNEL_NAMES.get(info.kernel_in_use, "?"),
                "kernel_ms_avg": k_ms, "kernel_ms_min": float(kt.min()) if kt.size else None,
                "kernel_launches_timed": int(kt.size), "algorithmic_bytes_per_launch": b_alg,
                "bytes_per_nnz": b_alg / nnz, "peak_source": peak_src,
                "roofline_gflops": 2.0 * nnz / (b_alg / (peak * 1e9)) / 1e9,
                "whole_step_GBps": b_alg / (ms_step * 1e6),
                "reference_getB_GBps": reference_getB(m, nnz, vb) / (ms_step * 1e6)}
    traffic_path = os.path.join(ROOT, "profiles", f"traffic_{name}.json")
    if os.path.exists(traffic_path):  # dram bytes per launch from the committed ncu --set full capture
        roofline["traffic"] = json.load(open(traffic_path)).get("dram_bytes_per_launch")
    out = dict(name=name, m=m, n=n, nnz=nnz, vb=vb, dtype=dtype, ms_step=ms_step, gflops=2.0 * nnz / (ms_step * 1e6),
               roofline=roofline, info=info, max_rel=max_rel, convert_ms=float(np.median(conv[1:])),
               convert_ms_all=conv, launches=info.launches_per_spmv, e2e=None)

    if full and not args.no_e2e:
        x_host = w["x"].cpu().pin_memory()
        y_host = torch.empty(m, dtype=dtype).pin_memory()

        def e2e_sync():
            assert A.spmv_host(1.0, x_host, y_host) == 0
        e2e_sync_ms = timed_loop(torch, e2e_sync, 3, 20, windows)   # one synchronous call per step, nothing overlapped
        rt = 1e-12 if vb == 8 else 1e-5
        assert torch.allclose(y_host.to(device), y, rtol=rt, atol=0)
        # the headline e2e: a stream of independent SpMVs through the pipelined host-buffer call; every
        # step still uploads its own x and downloads its own y (4 rotating pinned buffer pairs)
        nbuf = 4
        xs_host = [x_host] + [x_host.clone().pin_memory() for _ in range(nbuf - 1)]
        ys_host = [y_host] + [torch.empty(m, dtype=dtype).pin_memory() for _ in range(nbuf - 1)]
        e_batch = max(3, min(steps, 50))

        def e2e_batch():
            assert A.spmv_host_batch(1.0, [xs_host[i % nbuf] for i in range(e_batch)],
                                     [ys_host[i % nbuf] for i in range(e_batch)]) == 0
        ms_batch = timed_loop(torch, e2e_batch, 1, 2, windows)
        for yh in ys_host:   # carries of multi-tile rows are added with atomics: the last bits may differ between runs
            assert torch.allclose(yh.to(device), y, rtol=rt, atol=0), "host-buffer path disagrees with the device path"
        e_ms = ms_batch / e_batch
        out["e2e"] = {"value": 2.0 * nnz / (e_ms * 1e6), "unit": "GFLOP/s", "h2d_bytes_per_step": n * vb,
                      "d2h_bytes_per_step": m * vb, "ms_per_step": e_ms, "steps": 2 * e_batch,
                      "api": (f"csr5b200_spmv_host_batch: {e_batch} independent SpMVs per call, each with its own pinned x H2D "
                              "and y D2H, software-pipelined (upload k+1 | SpMV k | download k-1); CSR5 matrix resident"),
                      "ms_per_step_unpipelined_single_call": e2e_sync_ms}
    A.free()
    del w, y
    torch.cuda.empty_cache()
    return out


def workload_summary(r):
    """The part of a single-GPU result that goes under extra_workloads."""
    i = r["info"]
    return {"workload": WORKLOADS[r["name"]], "value": r["gflops"], "unit": "GFLOP/s", "ms_per_step": r["ms_step"],
            "dtype": "f64" if r["vb"] == 8 else "f32", "m": r["m"], "nnz": r["nnz"], "sigma": i.sigma,
            "num_packet": i.num_packet, "kernel": r["roofline"]["kernel"], "kernel_ms": r["roofline"]["kernel_ms_avg"],
            "roofline_frac": r["roofline"]["frac"], "achieved_GBps": r["roofline"]["achieved"],
            "algorithmic_bytes_per_launch": r["roofline"]["algorithmic_bytes_per_launch"],
            "traffic": r["roofline"]["traffic"], "csr_to_csr5_ms": r["convert_ms"],
            "parity_max_rel_err_vs_fp64_segment_sums": r["max_rel"], "parity": "pass",
            "hot_columns": i.hot_columns}


def main_single(args, torch, H, device):
    sampler_uuid = None
    try:
        sampler_uuid = str(torch.cuda.get_device_properties(0).uuid)
    except Exception:
        pass
    sampler = ClockSampler(device.index or 0, uuid=sampler_uuid)
    sampler.start()
    windows = []
    r = measure_single(torch, H, args.workload, args, device, windows, args.steps, full=True)
    extra = {}
    if args.workload == "c2" and not args.no_extra:
        for name in ("c3", "c4"):
            try:
                extra[name] = workload_summary(measure_single(torch, H, name, args, device, [], min(args.steps, 100), full=False))
            except Exception as e:   # never lose the headline line to a side measurement
                extra[name] = {"error": f"{type(e).__name__}: {e}"}
            log(f"[bench] extra workload {name}: {json.dumps(extra[name])[:300]}")
    sampler.stop()
    cpu = None
    if not args.no_cpu_baseline:
        c = cpu_reference_run(args.workload, 5, 30)
        cpu = {"value": c["value"], "unit": "GFLOP/s", "cores": c["cores"], "kind": c["kind"],
               "sample": c["sample"], "ms_per_spmv": c["ms"]}
    i, vb = r["info"], r["vb"]
    b_alg = r["roofline"]["algorithmic_bytes_per_launch"]
    emit({
        "metric": "FP64 SpMV GFLOPS" if vb == 8 else "FP32 SpMV GFLOPS",
        "value": r["gflops"], "unit": "GFLOP/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": r["ms_step"], "higher_is_better": True,
        "scaling": "strong" if args.workload in STRONG else "weak", "vs_baseline": None,
        "dtype": "f64" if vb == 8 else "f32", "data": "synthetic",
        "config": {
            "workload": WORKLOADS[args.workload], "m": r["m"], "n": r["n"], "nnz": r["nnz"], "sigma": i.sigma,
            "sigma_rule": "reference table (anonymouslib_cuda.h:297-313)" if args.sigma_rule == 0 else "measured on B200 (CSR5B200_OPT_SIGMA_RULE = 1)",
            "omega": 32, "tiles_per_gpu": i.p, "num_packet": i.num_packet, "values": "uniform (0,1], seed 42",
            "l2": f"inputs larger than L2 ({b_alg / 1e6:.0f} MB streamed per step vs 126 MB L2); no flush needed",
            "kernel": r["roofline"]["kernel"], "launches_per_step": r["launches"],
            "hot_columns": i.hot_columns, "hot_coverage": i.hot_coverage,
            "csr_to_csr5_ms": r["convert_ms"], "csr_to_csr5_in_spmvs": r["convert_ms"] / r["ms_step"],
            "csr_to_csr5_ms_first_call": r["convert_ms_all"][0], "csr_to_csr5_ms_samples": r["convert_ms_all"][1:],
            "parity_max_rel_err_vs_fp64_segment_sums": r["max_rel"],
        },
        "clocks": sampler.summary(windows),
        "e2e": r["e2e"],
        "gpu_launches": r["launches"] * args.steps,
        "roofline": r["roofline"],
        "cpu_baseline": cpu,
        "extra_workloads": extra or None,
        "multi_gpu": None,
    })


# ---------------------------------------------------------------------------------------------
# N > 1: one process per GPU
# ---------------------------------------------------------------------------------------------
def gather_rows(torch, dist, local, bounds, rank, device):
    """all-gather-v of per-rank row vectors (padded) -> the full vector on every rank."""
    sizes = [int(bounds[g + 1] - bounds[g]) for g in range(len(bounds) - 1)]
    mx = max(sizes)
    pad = torch.zeros(mx, dtype=local.dtype, device=device)
    pad[:local.numel()] = local
    tmp = torch.empty(len(sizes) * mx, dtype=local.dtype, device=device)
    dist.all_gather_into_tensor(tmp, pad)
    return torch.cat([tmp[g * mx:g * mx + s] for g, s in enumerate(sizes)])


def make_sharded(S, w, n, args, exch, transport=None, chunks=None, push_ctas=None):
    kw = dict(sigma=args.sigma, sigma_rule=args.sigma_rule)
    if exch == "nccl":
        return S.ShardedCsr5(w["bounds"], n, w["row_ptr"], w["col"], w["val"], mode="nccl", **kw)
    if exch in ("fused", "fused-unicast"):
        return S.ShardedCsr5(w["bounds"], n, w["row_ptr"], w["col"], w["val"], mode="fused",
                             multicast=None if exch == "fused" else False, scheme=args.scheme, **kw)
    return S.ShardedCsr5(w["bounds"], n, w["row_ptr"], w["col"], w["val"], mode="overlap",
                         transport=transport if transport is not None else args.transport,
                         chunks=chunks if chunks is not None else args.chunks,
                         push_ctas=push_ctas if push_ctas is not None else args.push_ctas,
                         push_threads=args.push_threads, **kw)


def measure_sharded(torch, dist, H, S, name, args, device, rank, world, windows, steps, full):
    """Workload `name` row-sharded over the ranks: parity of the gathered y on every rank, step time with the
    selected exchange, the same without exchange and with the NCCL all-gather."""
    w = build_workload(name, torch, device, rank, world)
    m, n, nnz, dtype = w["m"], w["n"], w["col"].numel(), w["dtype"]
    vb = 8 if dtype == torch.float64 else 4
    bounds = w["bounds"]
    m_total = int(bounds[-1])

    def sync_all():
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()

    def rank_max(v):
        t = torch.tensor([v], device=device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def all_ok(flag):
        t = torch.tensor([1 if flag else 0], device=device, dtype=torch.int32)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return bool(t.item())

    exch = args.exchange
    sh = None
    try:
        sh = make_sharded(S, w, n, args, exch)
        ok = True
    except Exception as e:   # e.g. symmetric memory unavailable on this box
        log(f"[bench] rank {rank}: exchange '{exch}' unavailable ({type(e).__name__}: {e}); falling back to nccl")
        ok = False
    if not all_ok(ok):
        if sh is not None:
            sh.free()
        exch = "nccl"
        sh = make_sharded(S, w, n, args, exch)
    A = sh.h
    assert sh.setX(w["x"]) == 0
    apply_tuning(A, H, args)
    A.warmup()
    torch.cuda.synchronize()
    conv = []
    for it in range(4):
        t0 = time.perf_counter()
        err = sh.asCSR5()
        torch.cuda.synchronize()
        conv.append((time.perf_counter() - t0) * 1e3)
        assert err == 0, A.error_string(err)
        if it < 3:
            assert A.asCSR() == 0

    # ---- parity: every rank's gathered y, element by element, against the all-gathered per-rank references ----
    y_full = sh.spmv(1.0)
    torch.cuda.synchronize()
    status = sh.exchange_status()
    A.asCSR()
    ref, scale = reference_rows(torch, w)
    ref_full = gather_rows(torch, dist, ref, bounds, rank, device)
    scale_full = gather_rows(torch, dist, scale, bounds, rank, device)
    max_rel = float(((y_full.double() - ref_full).abs() / scale_full.clamp_min(1e-300)).max())
    del ref, scale, ref_full, scale_full
    assert A.asCSR5() == 0
    tol = 1e-6 if vb == 8 else 1e-4
    good = all_ok(max_rel <= tol and status == 0)
    assert good, f"{name}: gathered y fails parity on some rank (rank {rank}: max rel err {max_rel}, exchange status {status})"
    max_rel = rank_max(max_rel)

    def step():
        sh.spmv(1.0)

    ms_step = rank_max(timed_loop(torch, step, max(args.warmup, 3), steps, windows, sync_all))
    assert all_ok(sh.exchange_status() == 0), "a device-side barrier timed out"
    info = A.info()
    default_transport, default_chunks = info.exchange_transport, info.exchange_chunks
    launches = info.launches_per_spmv
    t = torch.tensor([nnz], device=device, dtype=torch.int64)
    dist.all_reduce(t)
    total_nnz = int(t.item())

    k2 = max(3, min(steps, 100))
    ms_local = rank_max(timed_loop(torch, lambda: sh.spmv_local(1.0), 3, k2, None, sync_all))

    def nccl_step():
        sh.spmv_local(1.0)
        S.allgather_v(sh.y_full, bounds, rank)
    ms_nccl = rank_max(timed_loop(torch, nccl_step, 3, k2, None, sync_all))

    trace = None
    if args.trace_exchange and exch == "overlap":
        A.set_option(H.OPT_EXCHANGE_TRACE, 1)
        for _ in range(3):
            step()
        blocks, end = A.exchange_trace()
        A.set_option(H.OPT_EXCHANGE_TRACE, 0)
        mine = {"rank": rank, "rows": m, "blocks_ms_tiles_carry_shipped": blocks, "end_ms": end}
        allt = [None] * world
        dist.all_gather_object(allt, mine)
        trace = allt
    variants = {}
    if args.sweep_exchange and exch == "overlap":
        for spec in args.sweep_exchange.split("+"):
            tr, _, rest = spec.partition("/")
            ch, _, rest2 = rest.partition("/")
            ctas, _, thr = rest2.partition("/")
            if tr == "multicast" and not sh.has_multicast:
                variants[spec] = "no multicast address"
                continue
            sh.transport = H.TRANSPORT_NAMES[tr]
            sh.chunks = int(ch or 0)
            sh.push_ctas = int(ctas or 0)
            sh.push_threads = int(thr or 0)
            v = rank_max(timed_loop(torch, step, 3, k2, None, sync_all))
            st_ok = all_ok(sh.exchange_status() == 0)
            variants[spec] = v if st_ok else "barrier timeout"
        sh.transport, sh.chunks, sh.push_ctas = H.TRANSPORT_NAMES[args.transport], args.chunks, args.push_ctas
        sh.push_threads = args.push_threads

    if exch == "overlap":
        xi = A.info()   # transport / row blocks the (auto) rule resolved to in the last default step
        how = {1: "the copy engines", 2: "a push grid (unicast peer stores)", 3: "a push grid (NVSwitch multicast stores)",
               4: "the SpMV kernel itself (peer stores as tiles complete)", 5: "nobody"}.get(default_transport, "?")
        mode = (f"overlap: SpMV in {default_chunks} row block(s), finished blocks shipped by {how}; device-side flag "
                f"barrier; y double-buffered")
    elif exch == "nccl":
        mode = "nccl all-gather after the SpMV"
    else:
        mode = ("fused, " + {1: "rows stored to all GPUs by the SpMV kernels", 2: "coalesced push pass after the SpMV"}[sh.scheme]
                + (", NVSwitch multicast stores" if sh.multicast else ", unicast peer stores"))
    link_gbs = 770.0  # measured peer-copy bandwidth per direction per GPU (B200_PROFILING.md)
    in_bytes = (m_total - m) * vb
    multi = {"exchange": mode, "ms_per_step": ms_step, "ms_per_step_spmv_only_no_exchange": ms_local,
             "ms_per_step_spmv_then_nccl_allgather": ms_nccl,
             "nvlink_inbound_bytes_per_gpu_per_step": in_bytes, "nvlink_peak_GBps_per_direction": link_gbs,
             "nvlink_time_floor_ms": in_bytes / (link_gbs * 1e6),
             "step_floor_ms": max(ms_local, in_bytes / (link_gbs * 1e6)),
             "frac_of_step_floor": max(ms_local, in_bytes / (link_gbs * 1e6)) / ms_step,
             "exchange_variants_ms": variants or None, "exchange_trace": trace,
             "note": "every rank ends each step holding all of y: (N-1)/N of y must enter each GPU over NVLink per "
                     "step, which bounds the step from below next to the HBM stream"}
    out = dict(name=name, m=m, n=n, nnz=nnz, total_nnz=total_nnz, m_total=m_total, vb=vb, dtype=dtype, ms_step=ms_step,
               gflops=2.0 * total_nnz / (ms_step * 1e6), info=info, max_rel=max_rel, convert_ms=float(np.median(conv[1:])),
               convert_ms_all=conv, launches=launches, multi=multi, ms_local=ms_local, e2e=None, exch=exch,
               bounds=[int(b) for b in bounds])

    # ---- end to end, pipelined: every step uploads x and downloads y ---------------------------------------------
    if full and not args.no_e2e:
        out["e2e"] = e2e_sharded(torch, dist, sh, w, rank, world, device, n, m, m_total, vb, dtype, total_nnz,
                                 max(3, min(steps, 50)), windows, sync_all, rank_max)
    sh.free()
    del w, sh
    torch.cuda.empty_cache()
    return out


def e2e_sharded(torch, dist, sh, w, rank, world, device, n, m, m_total, vb, dtype, total_nnz, steps, windows, sync_all,
                rank_max):
    """Per step and per rank: pinned H2D of ITS 1/N slice of x, all-gather of x over NVLink (NCCL), the sharded SpMV
    with its y exchange, D2H of the rank's y rows -- software-pipelined over four streams with two buffers each, so
    upload k+1 | x all-gather k+1 | SpMV + y exchange k | download k-1 overlap (as csr5b200_spmv_host_batch does on
    one GPU)."""
    assert n % world == 0
    xs0, xs1 = rank * (n // world), (rank + 1) * (n // world)
    nbuf = 2
    x_host = [w["x"][xs0:xs1].cpu().pin_memory() for _ in range(nbuf)]
    y_host = [torch.empty(m, dtype=dtype).pin_memory() for _ in range(nbuf)]
    x_dev = [w["x"].clone() for _ in range(nbuf)]
    main = torch.cuda.current_stream()
    s_in, s_ag, s_out = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()
    e_in = [torch.cuda.Event() for _ in range(nbuf)]
    e_ag = [torch.cuda.Event() for _ in range(nbuf)]
    e_comp = [torch.cuda.Event() for _ in range(nbuf)]
    e_out = [torch.cuda.Event() for _ in range(nbuf)]
    state = {"k": 0}

    def one():
        k = state["k"]
        b = k % nbuf
        with torch.cuda.stream(s_in):
            if k >= nbuf:
                s_in.wait_event(e_comp[b])          # the SpMV that last read x_dev[b] is done
            x_dev[b][xs0:xs1].copy_(x_host[b], non_blocking=True)
            e_in[b].record(s_in)
        with torch.cuda.stream(s_ag):
            s_ag.wait_event(e_in[b])
            dist.all_gather_into_tensor(x_dev[b], x_dev[b][xs0:xs1])
            e_ag[b].record(s_ag)
        main.wait_event(e_ag[b])
        if k >= nbuf:
            main.wait_event(e_out[b])               # the download that last read this y buffer is done
        sh.setX(x_dev[b])
        sh.spmv(1.0)
        y_local = sh.y_local
        e_comp[b].record(main)
        with torch.cuda.stream(s_out):
            s_out.wait_event(e_comp[b])
            y_host[b].copy_(y_local, non_blocking=True)
            e_out[b].record(s_out)
        state["k"] = k + 1

    def sync_streams():
        for s in (s_in, s_ag, s_out):
            s.synchronize()
        sync_all()

    for _ in range(3):
        one()
    sync_streams()
    h0 = time.perf_counter()
    t0 = time.perf_counter()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(main)
    for _ in range(steps):
        one()
    main.wait_stream(s_out)
    ev1.record(main)
    sync_streams()
    windows.append((h0, time.perf_counter()))
    e_ms = rank_max(ev0.elapsed_time(ev1) / steps)
    sh.setX(w["x"])

    # the host link alone: the same uploads and downloads with nothing in between, both directions at once on all
    # ranks -- what this box's PCIe / host memory can carry, i.e. the floor of any end-to-end step
    def copies():
        b = state["k"] % nbuf
        with torch.cuda.stream(s_in):
            x_dev[b][xs0:xs1].copy_(x_host[b], non_blocking=True)
        with torch.cuda.stream(s_out):
            y_host[b].copy_(sh.y_local, non_blocking=True)
        state["k"] += 1
    for _ in range(2):
        copies()
    sync_streams()
    ev0.record(main)
    for s_ in (s_in, s_out):
        s_.wait_stream(main)
    for _ in range(steps):
        copies()
    main.wait_stream(s_in)
    main.wait_stream(s_out)
    ev1.record(main)
    sync_streams()
    copy_ms = rank_max(ev0.elapsed_time(ev1) / steps)
    rt = 1e-12 if vb == 8 else 1e-5
    for yh in y_host:
        assert torch.allclose(yh.to(device), sh.y_local, rtol=rt, atol=0), "host-buffer path disagrees with the device path"
    return {"value": 2.0 * total_nnz / (e_ms * 1e6), "unit": "GFLOP/s", "h2d_bytes_per_step": n * vb,
            "d2h_bytes_per_step": m_total * vb, "ms_per_step": e_ms, "steps": steps,
            "host_link_floor_ms_per_step": copy_ms,
            "host_link_GBps_both_directions_all_ranks": (n * vb + m_total * vb) / (copy_ms * 1e6),
            "host_link_note": "the same pinned uploads and downloads with no work in between, H2D and D2H concurrently on "
                              "all ranks: the floor this box's host link sets for any end-to-end step",
            "api": ("per rank and step: pinned H2D of its 1/N slice of x, NCCL all-gather of x, ShardedCsr5.spmv (SpMV + "
                    "overlapped y exchange), D2H of the rank's y rows; four streams, two buffers each, software-pipelined; "
                    "bytes are whole-job totals")}


def main_multi(args, torch, H, device, rank, world, local_rank):
    import torch.distributed as dist
    from benchmark_spmv_using_csr5_b200 import sharded as S
    numa = bind_to_gpu_numa(torch, local_rank) if not args.no_numa else {"bound": False, "skipped": True}
    dist.init_process_group("nccl", device_id=device)
    try:
        dev_uuid = str(torch.cuda.get_device_properties(local_rank).uuid)
    except Exception:
        dev_uuid = None
    sampler = ClockSampler(local_rank, uuid=dev_uuid)
    sampler.start()
    windows = []
    r = measure_sharded(torch, dist, H, S, args.workload, args, device, rank, world, windows, args.steps, full=True)

    c5 = None
    if args.workload == "c2" and not args.no_c5:
        # the strong-scaling claim of the north_star (R-MAT 25 over the GPUs of the box), inside the driver's line
        try:
            c5r = measure_sharded(torch, dist, H, S, "c5", args, device, rank, world, [], min(args.steps, 100), full=False)
            ms1 = None
            if rank == 0:
                one = measure_single(torch, H, "c5", args, device, [], min(args.steps, 50), full=False)
                ms1 = one["ms_step"]
                frac1 = one["roofline"]["frac"]
            t = torch.tensor([ms1 or 0.0, frac1 if rank == 0 else 0.0], device=device, dtype=torch.float64)
            dist.broadcast(t, 0)
            ms1, frac1 = float(t[0]), float(t[1])
            c5 = {"workload": WORKLOADS["c5"], "ms_1gpu": ms1, "roofline_frac_1gpu": frac1, "ms_N": c5r["ms_step"],
                  "speedup": ms1 / c5r["ms_step"], "gflops_N": c5r["gflops"], "nnz": c5r["total_nnz"],
                  "ms_N_spmv_only_no_exchange": c5r["ms_local"], "exchange": c5r["multi"]["exchange"],
                  "partition": f"contiguous row ranges minimising max(nnz, {args.row_cost:g} * rows) per shard "
                               "(a shard's SpMV costs its non-zeros, its share of the overlapped y exchange its rows)",
                  "rows_per_shard": [int(b - a) for a, b in zip(c5r["bounds"][:-1], c5r["bounds"][1:])],
                  "ms_N_spmv_then_nccl_allgather": c5r["multi"]["ms_per_step_spmv_then_nccl_allgather"],
                  "parity_max_rel_err": c5r["max_rel"], "parity": "pass (every rank's gathered y, element-wise)",
                  "exchange_variants_ms": c5r["multi"]["exchange_variants_ms"],
                  "exchange_trace": c5r["multi"]["exchange_trace"]}
        except Exception as e:
            c5 = {"error": f"{type(e).__name__}: {e}"}
            log(f"[bench] rank {rank}: c5 strong-scaling measurement failed: {c5['error']}")
    sampler.stop()
    if rank != 0:
        dist.destroy_process_group()
        return
    i, vb = r["info"], r["vb"]
    peak, peak_src = hbm_peak()
    b_alg = algorithmic_bytes(r["m"], r["n"], r["nnz"], vb)
    k_ms = r["ms_local"]
    roofline = {"bound": "hbm", "achieved": b_alg / (k_ms * 1e6), "peak": peak, "unit": "GB/s",
                "frac": b_alg / (k_ms * 1e6) / peak, "traffic": None, "kernel": KERNEL_NAMES.get(i.kernel_in_use, "?"),
                "kernel_ms_avg": k_ms, "algorithmic_bytes_per_launch": b_alg, "peak_source": peak_src,
                "note": "per GPU: this rank's SpMV (kernel + carry pass) timed without the exchange; the step is "
                        "bounded by NVLink, see multi_gpu"}
    r["multi"]["c5_strong"] = c5
    r["multi"]["numa"] = numa
    emit({
        "metric": "FP64 SpMV GFLOPS" if vb == 8 else "FP32 SpMV GFLOPS",
        "value": r["gflops"], "unit": "GFLOP/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": r["ms_step"], "higher_is_better": True,
        "scaling": "strong" if args.workload in STRONG else "weak", "vs_baseline": None,
        "dtype": "f64" if vb == 8 else "f32", "data": "synthetic",
        "config": {
            "workload": WORKLOADS[args.workload] + (f"; rank g owns rows [bounds[g], bounds[g+1]) of the {r['m_total']}-row "
                                                    f"matrix, x replicated, y concatenated on every rank each step ({r['multi']['exchange']})"),
            "m": r["m_total"], "n": r["n"], "nnz": r["total_nnz"], "sigma": i.sigma,
            "sigma_rule": "reference table (anonymouslib_cuda.h:297-313)" if args.sigma_rule == 0 else "measured on B200 (CSR5B200_OPT_SIGMA_RULE = 1)",
            "omega": 32, "tiles_per_gpu": i.p,
            "num_packet": i.num_packet, "values": "uniform (0,1], seed 42",
            "l2": f"inputs larger than L2 ({b_alg / 1e6:.0f} MB streamed per GPU per step vs 126 MB L2); no flush needed",
            "kernel": roofline["kernel"], "launches_per_step": r["launches"],
            "csr_to_csr5_ms": r["convert_ms"], "csr_to_csr5_ms_samples": r["convert_ms_all"],
            "parity_max_rel_err_vs_fp64_segment_sums": r["max_rel"],
            "parity": "every rank's gathered y compared element-wise with the all-gathered per-rank FP64 references",
            "rows_per_shard": [int(b - a) for a, b in zip(r["bounds"][:-1], r["bounds"][1:])],
            "partition_row_cost": args.row_cost if args.workload in STRONG else None,
        },
        "clocks": sampler.summary(windows),
        "e2e": r["e2e"],
        "gpu_launches": r["launches"] * args.steps,
        "roofline": roofline,
        "cpu_baseline": None,
        "extra_workloads": None,
        "multi_gpu": r["multi"],
    })
    dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--kernel", type=int, default=0, help="0 auto (= direct-load), 1 direct-load, 2 TMA-staged, 4 TMA-staged + x prefetch")
    ap.add_argument("--stages", type=int, default=0)
    ap.add_argument("--warps", type=int, default=0)
    ap.add_argument("--ctas-per-sm", type=int, default=0)
    ap.add_argument("--sigma", type=int, default=-1)
    ap.add_argument("--sigma-rule", type=int, default=1, help="0 = the reference's table (anonymouslib_cuda.h:297-313), 1 = the rule measured on B200")
    ap.add_argument("--hot", type=int, default=0, help="hot-column table: 0 off (default), -1 auto, K entries")
    ap.add_argument("--hot-threads", type=int, default=0)
    ap.add_argument("--wpb", type=int, default=0, help="tuning: warps per CTA of the direct kernel")
    ap.add_argument("--nch", type=int, default=0, help="tuning: register chunks per tile")
    ap.add_argument("--exchange", default="overlap", choices=["overlap", "fused", "fused-unicast", "nccl"],
                    help="N > 1: overlap = SpMV in row blocks with the finished blocks shipped meanwhile (default); fused = "
                         "first-generation in-kernel stores / push pass; nccl = all-gather after the SpMV")
    ap.add_argument("--transport", default="auto", choices=["auto", "ce", "push", "multicast", "inkernel", "none"])
    ap.add_argument("--chunks", type=int, default=0, help="overlap: row blocks per step (0 = default)")
    ap.add_argument("--push-ctas", type=int, default=0, help="overlap, SM transports: CTAs of the push grid (0 = default)")
    ap.add_argument("--sweep-exchange", default="", help="overlap: also time these variants, e.g. 'ce/8+push/8/32+multicast/8/32' (transport/chunks/push CTAs)")
    ap.add_argument("--push-threads", type=int, default=0, help="overlap, SM transports: threads per CTA of the push grid")
    ap.add_argument("--row-cost", type=float, default=8.0,
                    help="strong-scaling workloads (c3 / c5 on N > 1 GPUs): shards minimise max(nnz, cost * rows); 0 = balance "
                         "nnz alone.  8 bytes of y per row at the ~0.25 TB/s one GPU's multicast stream sustains against 12 "
                         "bytes per non-zero at half the HBM rate: a row costs about 8 non-zeros (DESIGN.md s6)")
    ap.add_argument("--trace-exchange", action="store_true", help="overlap: per-row-block timeline of one step on every rank")
    ap.add_argument("--scheme", type=int, default=0, help="fused modes: 0 auto, 1 in-kernel stores, 2 push pass")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="N = 1: skip the c3 / c4 side measurements")
    ap.add_argument("--no-c5", action="store_true", help="N > 1: skip the R-MAT 25 strong-scaling side measurement")
    ap.add_argument("--no-numa", action="store_true")
    args = ap.parse_args()
    capture_stdout()
    ROW_COST[0] = args.row_cost
    if args.impl == "reference":
        return run_reference_arm(args)
    args.warmup = max(args.warmup, 3)

    import torch
    from benchmark_spmv_using_csr5_b200 import handle as H

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        log(f"[bench] note: --gpus {args.gpus} but WORLD_SIZE {world}; using WORLD_SIZE")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world == 1:
        return main_single(args, torch, H, device)
    return main_multi(args, torch, H, device, rank, world, local_rank)


if __name__ == "__main__":
    main()
