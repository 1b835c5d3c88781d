"""2-GPU smoke of the symmetric-memory plumbing + scatter SpMV, with per-step logging (debug aid)."""
import os, sys, time
import numpy as np
import torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
logf = open(os.path.join(ROOT, "gpurun_out", f"debug_symm_r{rank}.log"), "w")
def log(*a):
    print(f"[{time.time():.3f}] r{rank}:", *a, file=logf, flush=True)
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
log("init pg")
dist.init_process_group("nccl", device_id=dev)
log("pg ok; can_access_peer", [torch.cuda.can_device_access_peer(lr, j) for j in range(world) if j != lr])
t = torch.ones(4, device=dev); dist.all_reduce(t); torch.cuda.synchronize(); log("allreduce ok", t.tolist())
import torch.distributed._symmetric_memory as symm_mem
log("symm backend", symm_mem.get_backend(dev) if hasattr(symm_mem, "get_backend") else "?")
buf = symm_mem.empty(1 << 20, dtype=torch.float64, device=dev)
log("symm empty ok", buf.data_ptr())
hdl = symm_mem.rendezvous(buf, dist.group.WORLD)
log("rendezvous ok", [hex(p) for p in hdl.buffer_ptrs], "multicast", hex(hdl.multicast_ptr) if hdl.multicast_ptr else None)
buf.fill_(rank + 1); torch.cuda.synchronize()
hdl.barrier(channel=0); torch.cuda.synchronize(); log("barrier ok")
peer = hdl.get_buffer((rank + 1) % world, (16,), torch.float64)
log("peer read", peer[:2].tolist())
hdl.barrier(channel=0); torch.cuda.synchronize()

import oracle
from benchmark_spmv_using_csr5_b200 import matrices as M, sharded as S
from tests.cases import small_cases
for name, A, sigma in small_cases()[:6]:
    val, x = M.values(A.nnz, A.n, "int")
    y_ref = oracle.csr_spmv(A.m, A.row_ptr, A.col, val, x)
    bounds = S.row_partition(A.row_ptr, world)
    rp, ci, v = S.shard_csr(A.row_ptr, A.col, val, bounds[rank], bounds[rank + 1])
    for mode in ("fused", "nccl"):
        log(name, mode, "build")
        sh = S.ShardedCsr5(bounds, A.n, torch.from_numpy(np.ascontiguousarray(rp)).cuda(),
                           torch.from_numpy(np.ascontiguousarray(ci)).cuda(),
                           torch.from_numpy(np.ascontiguousarray(v)).cuda(), mode=mode, sigma=sigma)
        sh.setX(torch.from_numpy(x).cuda())
        assert sh.asCSR5() == 0
        sh.y_full.fill_(float("nan")); torch.cuda.synchronize(); dist.barrier()
        log(name, mode, "spmv")
        y = sh.spmv(1.0); torch.cuda.synchronize()
        log(name, mode, "equal:", bool(np.array_equal(y.cpu().numpy(), y_ref)))
        dist.barrier(); sh.free()
log("done")
dist.destroy_process_group()
