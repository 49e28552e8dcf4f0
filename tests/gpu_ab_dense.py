"""A/B of the dense per-axis contraction out[o,i,c] = sum_j Q[i,j] in[o,j,c] (eigenbasis passes of the
fast-diagonalisation coarse solve): fp64 tensor-core kernel (poms_axis_dense_dmma, DMMA m8n8k4) against
the scalar gather kernel (poms_axis_gather with W = n), n = 35, 131, 259, 515, every axis of an n^3 array.
    python tests/gpu_ab_dense.py [sizes=35,131,259,515]
Under ncu (DMMA pipe utilisation):
    ncu --set full --clock-control none -k regex:axis_dense_dmma -c 3 -o gpurun_out/dmma python tests/gpu_ab_dense.py 131"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from poms_b200 import _lib
from poms_b200.mg import _AxisOp, _pitch

sizes = [int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "35,131,259,515").split(",")]
dev = torch.device("cuda", 0)
L = _lib.lib()
for n in sizes:
    rng = np.random.default_rng(n)
    Q = rng.standard_normal((n, n))
    ld = _pitch(n)
    x = torch.zeros((n, n, ld), dtype=torch.float64, device=dev)
    x[..., :n] = torch.as_tensor(rng.standard_normal((n, n, n)), device=dev)
    z = np.zeros(n, dtype=np.int32)
    op = _AxisOp(z, Q, n, dev, dense=True)
    ref = {0: np.einsum("ij,jbc->ibc", Q, x[..., :n].cpu().numpy()),
           1: np.einsum("ij,ajc->aic", Q, x[..., :n].cpu().numpy()),
           2: np.einsum("ij,abj->abi", Q, x[..., :n].cpu().numpy())}
    for axis in range(3):
        row = "n=%4d axis %d:" % (n, axis + 1)
        for mode in ("gather", "dmma"):
            os.environ["POMS_B200_DENSE"] = mode
            y = torch.zeros_like(x)
            reps = 3 if n >= 259 else 10
            for _ in range(2):
                op.apply(x, y, (n, n, n), ld, ld, axis)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                op.apply(x, y, (n, n, n), ld, ld, axis)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            err = np.abs(y[..., :n].cpu().numpy() - ref[axis]).max() / np.abs(ref[axis]).max()
            row += "  %s %9.4f ms %7.2f TFLOP/s err %.1e" % (mode, ms, 2.0 * n ** 4 / ms / 1e9, err)
        print(row, flush=True)
print("done")
