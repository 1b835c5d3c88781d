#!/usr/bin/env python
"""bench.py -- CSR5 SpMV on B200: GFLOPS and achieved HBM GB/s against the streaming roofline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c3|c4] [--impl ours|reference]

A step is one SpMV (y = A x) over the resident matrix, the unit the reference times in its
benchmark loop (CSR5_cuda/main.cu:93-106: NUM_RUN back-to-back spmv() between two events).
N = 1 workload = BASELINE.json configs[1]: banded 10M x 10M, 16 nnz/row, FP64.  N > 1: one process
per GPU (torchrun), weak scaling -- rank g owns the row range [g*m, (g+1)*m) of the N*m-row banded
matrix (its own CSR5 arrays), x is replicated, and every step ends with the all-gather of the y
segments over NCCL/NVLink.

One JSON line on stdout (rank 0).  `value` = whole-job GFLOPS with everything resident in HBM;
`e2e` = the same through the host-buffer C-ABI call (csr5b200_spmv_host: pinned x H2D + SpMV +
y D2H every step); `roofline` = algorithmic bytes / live CUDA-event duration of the main SpMV kernel
against the measured HBM copy bandwidth; `cpu_baseline` = the reference's own CSR5_avx2 backend
(oracle/_ref, compiled from /root/reference) on this box's host cores on a bounded sample.
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

WORKLOADS = {
    "c2": "banded 10M x 10M, 16 nnz/row, FP64 (BASELINE.json configs[1])",
    "c3": "R-MAT scale 22, edge factor 16, FP64 (BASELINE.json configs[2])",
    "c4": "27-pt Laplacian 320^3, FP32 (BASELINE.json configs[3])",
    "c5": "R-MAT scale 25, edge factor 16, FP64, row-range sharded by balanced nnz (BASELINE.json configs[4])",
}
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
        bounds = S.row_partition(rp, world)
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


def exact_check(torch, w, y):
    """max_i |y_i - ref_i| / sum_j |a_ij x_j| with ref = per-row FP64 sums evaluated on the device by
    an independent route (torch index_add_ of the FP64 products onto their row; col/val must be in CSR
    order).  For the all-positive C2/C3 inputs this is the plain relative error; for the signed
    Laplacian it is the usual row-wise backward-error normalisation."""
    m = w["row_ptr"].numel() - 1
    counts = (w["row_ptr"][1:] - w["row_ptr"][:-1]).long()
    rows = torch.repeat_interleave(torch.arange(m, device=counts.device), counts)
    prod = w["val"].double() * w["x"].double()[w["col"].long()]
    ref = torch.zeros(m, device=prod.device, dtype=torch.float64).index_add_(0, rows, prod)
    scale = torch.zeros(m, device=prod.device, dtype=torch.float64).index_add_(0, rows, prod.abs_())
    del prod, rows
    return float(((y.double() - ref).abs() / scale.clamp_min(1e-300)).max())


# ---------------------------------------------------------------------------------------------
# CPU baseline: the reference's CSR5_avx2 backend on the host cores (bounded sample)
# ---------------------------------------------------------------------------------------------
def cpu_reference_run(workload, warmup, runs, sample_rows=None):
    """Times oracle/_ref/libref_avx2.so (the reference's own CSR5_avx2, all host threads) on a bounded
    sample of `workload`.  Returns dict(value GFLOPS, ms, cores, kind, sample, nnz)."""
    import oracle
    from benchmark_spmv_using_csr5_b200 import matrices as M
    if workload == "c2":
        rows = sample_rows or 2_500_000
        A = M.banded(rows, 16)
        val, x = M.values(A.nnz, A.n, "real", np.float64, 42)
        sample = f"banded {rows} x {rows}, 16 nnz/row, FP64 ({A.nnz} nnz = 1/{10_000_000 // rows} of the workload's rows)"
    elif workload == "c3":
        A = M.rmat(18)
        val, x = M.values(A.nnz, A.n, "real", np.float64, 42)
        sample = f"R-MAT scale 18 ({A.nnz} nnz), FP64"
    elif workload == "c4":
        A, val = M.laplacian27(128)
        val = val.astype(np.float32)
        _, x = M.values(1, A.n, "real", np.float32, 42)
        ms, _y, threads = oracle.ref_csr_omp_bench_f32(A.m, A.row_ptr, A.col, val, x, 0, warmup, runs)
        return dict(value=2.0 * A.nnz / (ms * 1e6), ms=ms, cores=threads, kind="port", nnz=A.nnz,
                    sample=f"27-pt Laplacian 128^3 FP32 ({A.nnz} nnz), OpenMP scalar CSR loop "
                           "(the reference has no FP32 AVX2 path, README.md:36)")
    else:
        raise SystemExit(workload)
    if not oracle.ref_available():
        raise SystemExit("oracle/_ref/libref_avx2.so missing: run __graft_entry__.build() where /root/reference exists")
    ms, conv_ms, y, threads = oracle.ref_avx2_bench(A.m, A.n, A.row_ptr, A.col, val, x, 0, warmup, runs)
    y_ref = oracle.csr_spmv(A.m, A.row_ptr, A.col, val, x)
    ok = np.allclose(y, y_ref, rtol=1e-10, atol=0)
    return dict(value=2.0 * A.nnz / (ms * 1e6), ms=ms, cores=threads, kind="reference", nnz=A.nnz,
                sample=sample + f"; CSR5_avx2 sigma 16 omega 4; {warmup} warm-up + {runs} timed SpMVs"
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
# our arm
# ---------------------------------------------------------------------------------------------
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
    ap.add_argument("--hot", type=int, default=0, help="hot-column table: 0 off (default), -1 auto, K entries")
    ap.add_argument("--hot-threads", type=int, default=0)
    ap.add_argument("--scheme", type=int, default=0, help="N > 1 fused modes: 0 auto, 1 stores fused into the SpMV "
                    "kernels, 2 one coalesced push pass after the SpMV")
    ap.add_argument("--wpb", type=int, default=0, help="tuning: warps per CTA of the direct kernel")
    ap.add_argument("--nch", type=int, default=0, help="tuning: register chunks per tile")
    ap.add_argument("--exchange", default="fused", choices=["fused", "fused-unicast", "nccl"],
                    help="N > 1: y exchange fused into the SpMV kernels (peer stores) or NCCL all-gather after it")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    capture_stdout()
    if args.impl == "reference":
        return run_reference_arm(args)
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    from benchmark_spmv_using_csr5_b200 import handle as H

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        log(f"[bench] note: --gpus {args.gpus} but WORLD_SIZE {world}; using WORLD_SIZE")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)

    w = build_workload(args.workload, torch, device, rank, world)
    m, n, nnz, dtype = w["m"], w["n"], w["col"].numel(), w["dtype"]
    vb = 8 if dtype == torch.float64 else 4

    from benchmark_spmv_using_csr5_b200 import sharded as S
    bounds = w["bounds"]            # global row boundaries of the ranks' shards
    m_total = int(bounds[-1])
    r0, r1 = int(bounds[rank]), int(bounds[rank + 1])
    mode = args.exchange if world > 1 else "local"
    def make_handle(exch):
        return S.ShardedCsr5(bounds, n, w["row_ptr"], w["col"], w["val"], mode="nccl" if exch == "nccl" else "fused",
                             sigma=args.sigma, multicast=None if exch == "fused" else False, scheme=args.scheme)
    sh, ok = None, 1
    try:
        sh = make_handle(mode)
    except Exception as e:   # e.g. symmetric memory unavailable on this box
        log(f"[bench] rank {rank}: exchange '{mode}' unavailable ({type(e).__name__}: {e}); falling back to nccl")
        ok = 0
    if world > 1:
        t = torch.tensor([ok], device=device, dtype=torch.int32)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        ok = int(t.item())
    if not ok:
        if sh is not None:
            sh.free()
        mode = "nccl"
        sh = make_handle(mode)
    A = sh.h   # the ordinary single-GPU handle of this rank's rows
    assert sh.setX(w["x"]) == 0
    A.set_option(H.OPT_KERNEL, args.kernel)
    A.set_option(H.OPT_TMA_STAGES, args.stages)
    A.set_option(H.OPT_TMA_WARPS, args.warps)
    A.set_option(H.OPT_CTAS_PER_SM, args.ctas_per_sm)
    A.set_option(H.OPT_HOT_COLUMNS, args.hot)
    A.set_option(H.OPT_HOT_THREADS, args.hot_threads)
    A.set_option(H.OPT_EXCHANGE, args.scheme)
    A.set_option(H.OPT_DIRECT_WPB, args.wpb)
    A.set_option(H.OPT_DIRECT_NCH, args.nch)
    A.warmup()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    err = sh.asCSR5()
    torch.cuda.synchronize()
    convert_ms = (time.perf_counter() - t0) * 1e3
    assert err == 0, A.error_string(err)

    y_full = sh.y_full   # concatenated y (all ranks' segments)
    y = sh.y_local

    def step():
        sh.spmv(1.0)
        return 0

    # one checked SpMV (the reference checks its first call, main.cu:80-82, 360-384)
    assert step() == 0
    torch.cuda.synchronize()
    A.asCSR()  # the check needs col/val in CSR order; convert back afterwards
    max_rel = exact_check(torch, w, y)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    assert A.asCSR5() == 0
    torch.cuda.synchronize()
    convert_ms_cold, convert_ms = convert_ms, (time.perf_counter() - t0) * 1e3   # first call pays module loads
    tol = 1e-6 if vb == 8 else 1e-4
    ok = max_rel <= tol
    if world > 1:   # decide together: a rank that raised alone would leave the others in the next barrier
        t = torch.tensor([1 if ok else 0], device=device, dtype=torch.int32)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        ok = bool(t.item())
    assert ok, f"parity check failed (rank {rank}: max rel err {max_rel})"

    if world > 1:
        # every rank must hold the same concatenated y: compare a checksum of all segments
        cs = y_full.double().sum().reshape(1)
        allcs = [torch.empty_like(cs) for _ in range(world)]
        dist.all_gather(allcs, cs)
        assert all(torch.allclose(c, allcs[0], rtol=1e-12) for c in allcs), "ranks disagree on the gathered y"

    try:
        dev_uuid = str(torch.cuda.get_device_properties(local_rank).uuid)
    except Exception:
        dev_uuid = None
    sampler = ClockSampler(local_rank, uuid=dev_uuid)
    sampler.start()
    windows = []

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def time_loop(fn, warmup, steps):
        """W untimed + K timed calls of fn between two events on the current stream, bracketed by
        device sync + barrier; returns ms per step, max over ranks."""
        for _ in range(warmup):
            fn()
        sync_all()
        h0 = time.perf_counter()
        ev0.record()
        for _ in range(steps):
            fn()
        ev1.record()
        sync_all()
        windows.append((h0, time.perf_counter()))
        ms = ev0.elapsed_time(ev1) / steps
        if world > 1:
            t = torch.tensor([ms], device=device, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    # ---- device-resident timing -------------------------------------------------------------
    for _ in range(args.warmup):
        step()
    A.kernel_times_ms()  # drop
    A.set_option(H.OPT_KERNEL_TIMING, 1)
    ms_step = time_loop(step, 0, args.steps)
    kt = A.kernel_times_ms()
    A.set_option(H.OPT_KERNEL_TIMING, 0)
    info = A.info()
    launches_per_step = info.launches_per_spmv
    total_nnz = nnz
    if world > 1:
        t = torch.tensor([nnz], device=device, dtype=torch.int64)
        dist.all_reduce(t)
        total_nnz = int(t.item())
    gflops = 2.0 * total_nnz / (ms_step * 1e6)

    if world > 1 and mode != "nccl":
        mode = ("fused, " + {1: "rows stored to all GPUs by the SpMV kernels", 2: "coalesced push pass after the SpMV"}
                [sh.scheme] + (", NVSwitch multicast stores" if sh.multicast else ", unicast peer stores"))
    multi = None
    if world > 1:
        # the same step with the exchange done the other way, and without any exchange
        k2 = max(3, min(args.steps, 200))
        ms_local = time_loop(lambda: sh.spmv_local(1.0), 3, k2)

        def nccl_step():
            sh.spmv_local(1.0)
            S.allgather_v(y_full, bounds, rank)
        ms_nccl = time_loop(nccl_step, 3, k2)
        link_gbs = 770.0  # measured peer-copy bandwidth per direction per GPU (B200_PROFILING.md)
        in_bytes = (m_total - m) * vb
        t_link = in_bytes / (link_gbs * 1e6)
        multi = {"exchange": mode, "ms_per_step_spmv_only_no_exchange": ms_local,
                 "ms_per_step_spmv_then_nccl_allgather": ms_nccl, "ms_per_step_fused": ms_step
                 if mode.startswith("fused") else None,
                 "nvlink_inbound_bytes_per_gpu_per_step": in_bytes, "nvlink_peak_GBps_per_direction": link_gbs,
                 "nvlink_time_floor_ms": t_link,
                 "note": "every rank ends each step holding all of y: (N-1)*m*sizeof(VT) bytes must enter each "
                         "GPU over NVLink per step, which bounds the step from below next to the HBM stream"}

    # ---- end to end through the host-buffer C-ABI call -----------------------------------------
    e2e = None
    if not args.no_e2e:
        x_host = w["x"].cpu().pin_memory()
        y_host = torch.empty(m, dtype=dtype).pin_memory()
        e2e_sync_ms = None
        if world == 1:
            def e2e_step():
                assert A.spmv_host(1.0, x_host, y_host) == 0
            e2e_sync_ms = time_loop(e2e_step, 3, 20)   # one synchronous call per step, nothing overlapped
            assert torch.allclose(y_host.to(device), y, rtol=1e-12 if vb == 8 else 1e-5, atol=0)
            # the headline e2e: a stream of independent SpMVs through the pipelined host-buffer call; every
            # step still uploads its own x and downloads its own y (4 rotating pinned buffer pairs)
            nbuf = 4
            xs_host = [x_host] + [x_host.clone().pin_memory() for _ in range(nbuf - 1)]
            ys_host = [y_host] + [torch.empty(m, dtype=dtype).pin_memory() for _ in range(nbuf - 1)]
            e_batch = max(3, min(args.steps, 50))

            def e2e_step():
                assert A.spmv_host_batch(1.0, [xs_host[i % nbuf] for i in range(e_batch)],
                                         [ys_host[i % nbuf] for i in range(e_batch)]) == 0
            ms_batch = time_loop(e2e_step, 1, 2)
            api = (f"csr5b200_spmv_host_batch: {e_batch} independent SpMVs per call, each with its own pinned x H2D and "
                   "y D2H, software-pipelined (upload k+1 | SpMV k | download k-1); CSR5 matrix resident")
        else:
            x_dev = w["x"]
            xs0, xs1 = rank * (n // world), (rank + 1) * (n // world)   # rank g uploads its 1/N slice of x
            x_slice_dev, x_slice_host = x_dev[xs0:xs1], x_host[xs0:xs1]
            assert n % world == 0

            def e2e_step():   # every rank: upload ITS slice of x over PCIe, replicate x over NVLink (NCCL
                #               all-gather), sharded SpMV + fused y exchange, download its rows of y
                x_slice_dev.copy_(x_slice_host, non_blocking=True)
                dist.all_gather_into_tensor(x_dev, x_slice_dev)
                sh.spmv(1.0)
                y_host.copy_(y, non_blocking=True)
            api = ("per rank: pinned H2D of its 1/N slice of x, NCCL all-gather of x, ShardedCsr5.spmv (SpMV + y "
                   "exchange), D2H of the rank's y rows; bytes are whole-job totals")
        if world == 1:
            e_steps, e_ms = 2 * e_batch, ms_batch / e_batch
            for yh in ys_host:
                assert torch.allclose(yh.to(device), y, rtol=1e-12 if vb == 8 else 1e-5, atol=0)
        else:
            e_steps = max(3, min(args.steps, 50))
            e_ms = time_loop(e2e_step, 3, e_steps)
        torch.cuda.synchronize()
        # carries of multi-tile rows are added with atomics: the last bits may differ between runs
        assert torch.allclose(y_host.to(device), y, rtol=1e-12 if vb == 8 else 1e-5, atol=0), \
            "host-buffer path disagrees with the device path"
        e2e = {"value": 2.0 * total_nnz / (e_ms * 1e6), "unit": "GFLOP/s", "h2d_bytes_per_step": n * vb,
               "d2h_bytes_per_step": m_total * vb, "ms_per_step": e_ms, "steps": e_steps, "api": api,
               "ms_per_step_unpipelined_single_call": e2e_sync_ms}
    sampler.stop()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel ------------------------------------------------------
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy)"
    else:
        peak, peak_src = 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"
    b_alg = algorithmic_bytes(m, n, nnz, vb)
    k_ms = float(kt.mean()) if kt.size else ms_step
    achieved = b_alg / (k_ms * 1e6)
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": None, "kernel": {1: "spmv_direct_kernel", 2: "spmv_tma_kernel", 3: "spmv_hot_kernel",
                                                   4: "spmv_tma_kernel<prefetch>"}.get(info.kernel_in_use, "?"),
                "kernel_ms_avg": k_ms, "kernel_ms_min": float(kt.min()) if kt.size else None,
                "kernel_launches_timed": int(kt.size), "algorithmic_bytes_per_launch": b_alg,
                "bytes_per_nnz": b_alg / nnz, "peak_source": peak_src,
                "roofline_gflops": 2.0 * nnz / (b_alg / (peak * 1e9)) / 1e9,
                "whole_step_GBps": b_alg / (ms_step * 1e6),
                "reference_getB_GBps": reference_getB(m, nnz, vb) / (ms_step * 1e6)}
    traffic_path = os.path.join(ROOT, "profiles", f"traffic_{args.workload}.json")
    if os.path.exists(traffic_path):  # dram bytes per launch from the committed ncu --set full capture
        roofline["traffic"] = json.load(open(traffic_path)).get("dram_bytes_per_launch")

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        r = cpu_reference_run(args.workload, 5, 30)
        cpu = {"value": r["value"], "unit": "GFLOP/s", "cores": r["cores"], "kind": r["kind"],
               "sample": r["sample"], "ms_per_spmv": r["ms"]}

    out = {
        "metric": "FP64 SpMV GFLOPS" if vb == 8 else "FP32 SpMV GFLOPS",
        "value": gflops, "unit": "GFLOP/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong" if args.workload in STRONG else "weak", "vs_baseline": None,
        "dtype": "f64" if vb == 8 else "f32", "data": "synthetic",
        "config": {
            "workload": WORKLOADS[args.workload] + (f"; rank g owns rows [bounds[g], bounds[g+1]) of the {m_total}-row matrix, "
                                                    f"x replicated, y concatenated on every rank each step ({mode})" if world > 1 else ""),
            "m": m_total, "n": n, "nnz": total_nnz, "sigma": info.sigma, "omega": 32, "tiles_per_gpu": info.p,
            "num_packet": info.num_packet, "values": "uniform (0,1], seed 42", "l2": "inputs larger than L2 "
            f"({b_alg / 1e6:.0f} MB streamed per step vs 126 MB L2); no flush needed",
            "kernel": roofline["kernel"], "launches_per_step": launches_per_step,
            "hot_columns": info.hot_columns, "hot_coverage": info.hot_coverage,
            "csr_to_csr5_ms": convert_ms, "csr_to_csr5_in_spmvs": convert_ms / ms_step,
            "csr_to_csr5_ms_first_call": convert_ms_cold,
            "parity_max_rel_err_vs_fp64_segment_sums": max_rel,
        },
        "clocks": sampler.summary(windows),
        "e2e": e2e,
        "gpu_launches": launches_per_step * args.steps,
        "roofline": roofline,
        "cpu_baseline": cpu,
        "multi_gpu": multi,
    }
    emit(out)
    A.free()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
