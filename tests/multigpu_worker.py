"""Worker of tests/test_gpu_sharded.py, launched with torch.distributed.run (one process per GPU):
row-range sharded SpMV in both exchange modes against the oracle's y of the whole matrix."""
import os
import sys

os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")   # one hardware queue per stream (device-side barriers)

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    import oracle
    from benchmark_spmv_using_csr5_b200 import matrices as M
    from benchmark_spmv_using_csr5_b200 import sharded as S
    from tests.cases import small_cases
    bad = []
    for name, A, sigma in small_cases():
        if A.nnz == 0:
            continue
        for dt, tdt in ((np.float64, torch.float64), (np.float32, torch.float32)):
            val, x = M.values(A.nnz, A.n, "int", dt)
            y_ref = oracle.csr_spmv(A.m, A.row_ptr, A.col, val, x)
            bounds = S.row_partition(A.row_ptr, world)
            rp, ci, v = S.shard_csr(A.row_ptr, A.col, val, bounds[rank], bounds[rank + 1])
            for mode in ("overlap-ce", "overlap-push", "overlap-multicast", "overlap-inkernel",
                         "fused-multicast-1", "fused-unicast-1", "fused-multicast-2", "fused-unicast-2",
                         "fused-multicast-0", "nccl"):
                dev = [torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in (rp, ci, v)]
                if mode.startswith("overlap"):
                    sh = S.ShardedCsr5(bounds, A.n, *dev, mode="overlap", sigma=sigma, transport=mode.split("-")[1],
                                       chunks=3, push_ctas=8, timeout_ms=8000)
                    if mode == "overlap-multicast" and not sh.has_multicast:
                        sh.free()
                        continue
                else:
                    sh = S.ShardedCsr5(bounds, A.n, *dev, mode="nccl" if mode == "nccl" else "fused", sigma=sigma,
                                       multicast="multicast" in mode, scheme=int(mode[-1]) if mode != "nccl" else 0)
                sh.setX(torch.from_numpy(x).cuda())
                assert sh.asCSR5() == 0
                torch.cuda.synchronize()
                dist.barrier()
                for _ in range(3):   # repeated calls stay exact (no accumulation across calls; y buffers alternate)
                    y = sh.spmv(1.0)
                torch.cuda.synchronize()
                if sh.exchange_status() != 0 or not np.array_equal(y.cpu().numpy(), y_ref):
                    bad.append((name, dt.__name__, mode))
                if mode in ("overlap-ce", "overlap-push", "nccl"):
                    y_prev = y.clone()
                    y2 = sh.spmv(2.0, -1.0)   # beta refers to the y of the previous step
                    torch.cuda.synchronize()
                    if not np.array_equal(y2.cpu().numpy(), 2.0 * y_ref - y_prev.cpu().numpy()):
                        bad.append((name, dt.__name__, mode + "+beta"))
                dist.barrier()
                sh.free()
    # y -> x feedback across the GPUs: three steps of x <- A x on a square matrix
    A = M.banded(4096, 16)
    rng = np.random.default_rng(3)
    val = rng.integers(0, 3, size=A.nnz).astype(np.float64)
    x0 = rng.integers(0, 3, size=A.n).astype(np.float64)
    want = x0
    for _ in range(3):
        want = oracle.csr_spmv(A.m, A.row_ptr, A.col, val, want)
    bounds = S.row_partition(A.row_ptr, world)
    rp, ci, v = S.shard_csr(A.row_ptr, A.col, val, bounds[rank], bounds[rank + 1])
    for mode, tr in (("overlap", "ce"), ("overlap", "push"), ("nccl", "auto")):
        dev = [torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in (rp, ci, v)]
        sh = S.ShardedCsr5(bounds, A.n, *dev, mode=mode, transport=tr, chunks=4, timeout_ms=8000)
        sh.setX(torch.from_numpy(x0).cuda())
        assert sh.asCSR5() == 0
        y = sh.iterate(3)
        torch.cuda.synchronize()
        if sh.exchange_status() != 0 or not np.array_equal(y.cpu().numpy(), want):
            bad.append(("iterate", mode, tr))
        want4 = oracle.csr_spmv(A.m, A.row_ptr, A.col, val, want)
        for _ in range(3):   # x = the last iterate, in a buffer of its own: further steps must not overwrite it
            y = sh.spmv(1.0)
            torch.cuda.synchronize()
            if not np.array_equal(y.cpu().numpy(), want4):
                bad.append(("iterate+spmv", mode, tr))
        dist.barrier()
        sh.free()
    print(f"rank {rank}: {'OK' if not bad else 'MISMATCH ' + str(bad)}", flush=True)
    dist.destroy_process_group()
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
