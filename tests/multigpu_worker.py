"""Worker of tests/test_gpu_sharded.py, launched with torch.distributed.run (one process per GPU):
row-range sharded SpMV in both exchange modes against the oracle's y of the whole matrix."""
import os
import sys

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
            for mode in ("fused-multicast-1", "fused-unicast-1", "fused-multicast-2", "fused-unicast-2",
                         "fused-multicast-0", "nccl"):
                sh = S.ShardedCsr5(bounds, A.n, torch.from_numpy(np.ascontiguousarray(rp)).cuda(),
                                   torch.from_numpy(np.ascontiguousarray(ci)).cuda(),
                                   torch.from_numpy(np.ascontiguousarray(v)).cuda(),
                                   mode="nccl" if mode == "nccl" else "fused", sigma=sigma,
                                   multicast="multicast" in mode, scheme=int(mode[-1]) if mode != "nccl" else 0)
                sh.setX(torch.from_numpy(x).cuda())
                assert sh.asCSR5() == 0
                sh.y_full.fill_(float("nan"))
                torch.cuda.synchronize()
                dist.barrier()
                for _ in range(2):   # repeated calls stay exact (no accumulation across calls)
                    y = sh.spmv(1.0)
                torch.cuda.synchronize()
                if not np.array_equal(y.cpu().numpy(), y_ref):
                    bad.append((name, dt.__name__, mode))
                dist.barrier()
                sh.free()
    print(f"rank {rank}: {'OK' if not bad else 'MISMATCH ' + str(bad)}", flush=True)
    dist.destroy_process_group()
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
