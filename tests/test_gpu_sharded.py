"""Multi-GPU parity (needs >= 2 GPUs; skipped on a single-GPU box): the row-range sharded SpMV with
the y exchange fused into the kernels (peer stores over NVLink) and with the NCCL all-gather, against
the oracle's y of the whole matrix.  Launched through torch.distributed.run under a hard timeout so a
failing rank can never leave the others waiting in a collective."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_sharded_spmv_matches_oracle():
    import torch
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs >= 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(ROOT, "tests", "multigpu_worker.py")]
    try:
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=420, cwd=ROOT, start_new_session=True)
    except subprocess.TimeoutExpired as e:
        pytest.fail("sharded worker timed out:\n" + str(e.stdout)[-2000:])
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count(": OK") == world, r.stdout[-3000:]
