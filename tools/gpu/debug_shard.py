"""2-GPU step-by-step smoke of the sharded SpMV modes with per-step logging (debug aid)."""
import os, sys, time
import numpy as np
import torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
logf = open(os.path.join(ROOT, "gpurun_out", f"debug_shard_r{rank}.log"), "w")
def log(*a):
    print(f"[{time.time():.3f}] r{rank}:", *a, file=logf, flush=True)
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
import oracle
from benchmark_spmv_using_csr5_b200 import matrices as M, sharded as S
from tests.cases import small_cases
modes = sys.argv[1].split(",")
cases = small_cases()
cases = [cases[0], cases[4], cases[15]] + [("banded_big", M.banded(400000, 16), -1)]
for name, A, sigma in cases:
    val, x = M.values(A.nnz, A.n, "int")
    y_ref = oracle.csr_spmv(A.m, A.row_ptr, A.col, val, x)
    bounds = S.row_partition(A.row_ptr, world)
    rp, ci, v = S.shard_csr(A.row_ptr, A.col, val, bounds[rank], bounds[rank + 1])
    for mode in modes:
        log(name, mode, "build")
        sh = S.ShardedCsr5(bounds, A.n, torch.from_numpy(np.ascontiguousarray(rp)).cuda(),
                           torch.from_numpy(np.ascontiguousarray(ci)).cuda(),
                           torch.from_numpy(np.ascontiguousarray(v)).cuda(),
                           mode="nccl" if mode == "nccl" else "fused", sigma=sigma, multicast=(mode == "mc"))
        sh.setX(torch.from_numpy(x).cuda())
        assert sh.asCSR5() == 0
        sh.y_full.fill_(float("nan")); torch.cuda.synchronize(); dist.barrier()
        log(name, mode, "spmv enqueue; multicast =", sh.multicast)
        if sh._dst is not None:
            sh._resolve_exchange()
        err = sh.h.spmv_scatter(1.0, sh.y_local, sh._dst, len(sh._dst), sh.multicast) if sh._dst is not None else sh.h.spmv(1.0, sh.y_local)
        log(name, mode, "enqueued err", err)
        torch.cuda.synchronize()
        log(name, mode, "local sync ok; local rows equal:", bool(np.array_equal(sh.y_local.cpu().numpy(), y_ref[bounds[rank]:bounds[rank+1]])))
        dist.barrier(); torch.cuda.synchronize()
        log(name, mode, "full equal:", bool(np.array_equal(sh.y_full.cpu().numpy(), y_ref)) if mode != "nccl" else "n/a")
        dist.barrier(); sh.free()
log("done")
dist.destroy_process_group()
