"""Multi-GPU parity (needs >= 2 GPUs; skipped on a single-GPU box): the row-range sharded SpMV with
the y exchange fused into the kernels (peer stores over NVLink) and with the NCCL all-gather, against
the oracle's y of the whole matrix."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, result_dir):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import oracle
        from benchmark_spmv_using_csr5_b200 import matrices as M
        from benchmark_spmv_using_csr5_b200 import sharded as S
        from tests.cases import small_cases
        ok = True
        for name, A, sigma in small_cases():
            if A.nnz == 0:
                continue
            val, x = M.values(A.nnz, A.n, "int")
            y_ref = oracle.csr_spmv(A.m, A.row_ptr, A.col, val, x)
            bounds = S.row_partition(A.row_ptr, world)
            rp, ci, v = S.shard_csr(A.row_ptr, A.col, val, bounds[rank], bounds[rank + 1])
            for mode in ("fused", "nccl"):
                sh = S.ShardedCsr5(bounds, A.n, torch.from_numpy(np.ascontiguousarray(rp)).cuda(),
                                   torch.from_numpy(np.ascontiguousarray(ci)).cuda(),
                                   torch.from_numpy(np.ascontiguousarray(v)).cuda(), mode=mode, sigma=sigma)
                sh.setX(torch.from_numpy(x).cuda())
                assert sh.asCSR5() == 0
                sh.y_full.fill_(float("nan"))
                dist.barrier()
                for _ in range(2):   # repeated calls stay exact (no accumulation across calls)
                    y = sh.spmv(1.0)
                torch.cuda.synchronize()
                good = np.array_equal(y.cpu().numpy(), y_ref)
                if not good:
                    print(f"rank {rank}: {name} mode {mode} MISMATCH", flush=True)
                ok = ok and good
                dist.barrier()
                sh.free()
        open(os.path.join(result_dir, f"r{rank}"), "w").write("ok" if ok else "FAIL")
    finally:
        dist.destroy_process_group()


def test_sharded_spmv_matches_oracle(tmp_path):
    import torch
    import torch.multiprocessing as mp
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs >= 2 GPUs")
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert open(tmp_path / f"r{r}").read() == "ok"
