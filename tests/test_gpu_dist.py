"""Multi-GPU parity as a pytest test: runs tests/gpu_dist_check.py under torchrun on every GPU of
the box (2, 4 or 8; skipped below 2).  The script compares the slab-partitioned path (peer-store and
NCCL halo exchange, SPIKE solve, distributed transfers, gathered coarse levels, MG-PCG) with the CPU
oracle on the same global problems: identical iteration counts, 1e-8 on the solutions."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT


@pytest.mark.gpu
@pytest.mark.parametrize("v2_min", [None, "0"])
def test_slab_partitioned_path_matches_oracle(v2_min):
    """v2_min = "0": the round-2 one-pass transfer kernels (poms_transfer3d_v2.cu) also on the small
    grids of the script (by default they start at 1e6 / 6e6 fine points per rank)."""
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs (found %d)" % n)
    n = 8 if n >= 8 else 4 if n >= 4 else 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n),
           "--master-addr", "127.0.0.1", "--master-port", "29533" if v2_min is None else "29534",
           os.path.join(ROOT, "tests", "gpu_dist_check.py")]
    env = dict(os.environ)
    if v2_min is not None:
        env["POMS_B200_TRANSFER_V2_MIN"] = v2_min
    r = subprocess.run(cmd, cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True,
                       timeout=1500, env=env)
    print(r.stdout[-6000:])
    assert r.returncode == 0, r.stdout[-3000:]
    assert "ALL OK" in r.stdout
