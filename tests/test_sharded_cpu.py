"""CPU tests of the multi-GPU host logic: balanced row-range partitioning, shard
slicing/rebasing and the y all-gather(-v), run with world_size 2 and 3 on the gloo backend.  The
per-shard SpMV is done by the ORACLE here (tests may call it; the product's shards run the CUDA
handle) -- what is under test is that the concatenation of per-shard results equals the global y."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from benchmark_spmv_using_csr5_b200 import matrices as M
from benchmark_spmv_using_csr5_b200 import sharded as S
from tests.cases import small_cases


def test_row_partition_rules():
    for name, A, _ in small_cases():
        for G in (1, 2, 3, 8):
            b = S.row_partition(A.row_ptr, G)
            assert b[0] == 0 and b[-1] == A.m and np.all(np.diff(b) >= 0), name
            bt = S.row_partition(torch.from_numpy(A.row_ptr), G)
            assert np.array_equal(b, bt), name
            # interior boundary g = the last row whose first nnz index is <= g*nnz/G
            for g in range(1, G):
                tgt = g * A.nnz // G
                r = b[g]
                assert A.row_ptr[r] <= tgt, name
                if r < A.m:
                    assert A.row_ptr[r + 1] > tgt or r == b[g + 1], name
    # balanced to within the longest row
    A = M.rmat(12)
    b = S.row_partition(A.row_ptr, 8)
    per = np.diff(A.row_ptr[b])
    assert per.max() - per.min() <= 2 * np.diff(A.row_ptr).max()


def test_shard_csr_rebases():
    A = M.example_c1()
    val, _ = M.values(A.nnz, A.n, "int")
    b = S.row_partition(A.row_ptr, 3)
    tot = 0
    for g in range(3):
        rp, ci, v = S.shard_csr(A.row_ptr, A.col, val, b[g], b[g + 1])
        assert rp[0] == 0 and len(rp) == b[g + 1] - b[g] + 1 and rp[-1] == len(ci) == len(v)
        assert np.array_equal(ci, A.col[A.row_ptr[b[g]]:A.row_ptr[b[g + 1]]])
        tot += len(ci)
    assert tot == A.nnz


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, case_idx, result_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import oracle
        name, A, sigma = small_cases()[case_idx]
        val, x = M.values(A.nnz, A.n, "int")
        bounds = S.row_partition(A.row_ptr, world)
        rp, ci, v = S.shard_csr(A.row_ptr, A.col, val, bounds[rank], bounds[rank + 1])
        m_loc = int(bounds[rank + 1] - bounds[rank])
        y_loc = oracle.csr5_spmv(m_loc, A.n, np.ascontiguousarray(rp), np.ascontiguousarray(ci),
                                 np.ascontiguousarray(v), x, sigma) if len(ci) else np.zeros(m_loc)
        y_full = torch.full((A.m,), float("nan"), dtype=torch.float64)
        y_full[int(bounds[rank]):int(bounds[rank + 1])] = torch.from_numpy(y_loc)
        S.allgather_v(y_full, bounds, rank)
        y_ref = oracle.csr_spmv(A.m, A.row_ptr, A.col, val, x)
        ok = np.array_equal(y_full.numpy(), y_ref)
        open(os.path.join(result_dir, f"r{rank}"), "w").write("ok" if ok else "FAIL")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("case_idx", [4, 5, 15])  # empty rows + long row, example, hub row
def test_sharded_allgather_gloo(tmp_path, world, case_idx):
    port = _free_port()
    mp.spawn(_worker, args=(world, port, case_idx, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert open(tmp_path / f"r{r}").read() == "ok"


def test_row_partition_with_row_cost():
    """max(nnz, cost * rows) minimisation: the nnz rule where rows do not bind, fewer rows for the row-heavy shard where
    they do; never worse than the nnz rule under its own objective."""
    from benchmark_spmv_using_csr5_b200 import matrices as M
    from benchmark_spmv_using_csr5_b200 import sharded as S
    import numpy as np
    A = M.rmat(14)

    def objective(b, cost):
        return max(np.diff(A.row_ptr[b].astype(np.int64)).max(), cost * np.diff(b).max())

    for parts in (2, 4, 8):
        b0 = S.row_partition(A.row_ptr, parts)
        assert np.array_equal(b0, S.row_partition(A.row_ptr, parts, row_cost=0.0))
        b = S.row_partition(A.row_ptr, parts, row_cost=7.0)
        assert len(b) == parts + 1 and b[0] == 0 and b[-1] == A.m and np.all(np.diff(b) >= 0)
        assert objective(b, 7.0) <= objective(b0, 7.0)
    b0, b8 = S.row_partition(A.row_ptr, 8), S.row_partition(A.row_ptr, 8, row_cost=7.0)
    assert np.diff(b8).max() < np.diff(b0).max()            # the row-heaviest shard sheds rows at 8 shards
    try:
        import torch
        assert np.array_equal(S.row_partition(torch.from_numpy(A.row_ptr), 8, row_cost=7.0), b8)
    except ImportError:
        pass
    B = M.banded(10000, 16)                                 # uniform rows: the nnz rule either way
    assert np.array_equal(S.row_partition(B.row_ptr, 4, row_cost=7.0), S.row_partition(B.row_ptr, 4))
