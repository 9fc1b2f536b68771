"""Multi-GPU parity (needs >= 2 GPUs; skipped on a single-GPU box): tools/dist_check.py under torchrun, every rank running its
part of the mesh, compared on rank 0 with the single-GPU solver (1e-11).  Data planes: the default (peer-to-peer: every kernel
delivers the rows it produces straight into the other ranks' arrays and synchronises through epoch flags), the same plane with the
persistent visit kernel (MGCFD_VISIT=1: its grid barriers double as the halo exchange) and NCCL send/recv (MGCFD_NO_P2P=1)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("nranks", [2, 4])
@pytest.mark.parametrize("plane", ["p2p-visit", "p2p-stage", "nccl"])
def test_distributed_matches_single_gpu(nranks, plane):
    import torch
    if torch.cuda.device_count() < nranks:
        pytest.skip(f"needs {nranks} GPUs")
    port = 29600 + nranks + {"p2p-visit": 0, "p2p-stage": 10, "nccl": 20}[plane]
    env = dict(os.environ)
    if plane == "p2p-visit":
        env["MGCFD_VISIT"] = "1"
    if plane == "nccl":
        env["MGCFD_NO_P2P"] = "1"
    if plane == "p2p-stage":
        env["MGCFD_GUARD"] = "1"          # the default plane also with guard zones around every device array of every rank
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nranks}", "--master-addr", "127.0.0.1",
                        "--master-port", str(port), os.path.join(ROOT, "tools", "dist_check.py")], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0 and "dist_check PASS" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
