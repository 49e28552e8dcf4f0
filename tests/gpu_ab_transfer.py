"""A/B timing of the grid transfers: fused one-pass kernels vs per-axis gathers -- run on the GPU box:
    python tests/gpu_ab_transfer.py [N] [p] [Nmin]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from poms_b200 import bsplines as bs
from poms_b200.mg import Transfer
from poms_b200.stencil import StencilVectorSpace, StencilVector

N = int(sys.argv[1]) if len(sys.argv) > 1 else 512
p = int(sys.argv[2]) if len(sys.argv) > 2 else 3
Nmin = int(sys.argv[3]) if len(sys.argv) > 3 else 64
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(0)
while N >= Nmin:
    Tf = [bs.make_open_knots(p, N + p)] * 3
    Tc = [bs.make_open_knots(p, N // 2 + p)] * 3
    Vf = StencilVectorSpace([N + p] * 3, [p] * 3, [False] * 3, device=dev)
    Vc = StencilVectorSpace([N // 2 + p] * 3, [p] * 3, [False] * 3, device=dev)
    tr = Transfer(Tc, Tf, p, dev)
    rf, xf, ec = StencilVector(Vf), StencilVector(Vf), StencilVector(Vc)
    rf.data.copy_(torch.randn(Vf.npts, generator=g, dtype=torch.float64, device=dev))
    ec.data.copy_(torch.randn(Vc.npts, generator=g, dtype=torch.float64, device=dev))
    res = {}
    for fused in (False, True, "v2"):
        tr.fused, tr.fused_max = fused is True, (10 ** 12 if fused is True else -1)
        tr.fused_v2, tr.v2_min = fused == "v2", 0
        xf.flat.zero_()
        rc = tr.restrict(rf, Vc)
        tr.prolong_add(ec, xf)
        res[fused] = (rc.data.clone(), xf.data.clone())
        for name, fn in (("restrict", lambda: tr.restrict(rf, Vc)), ("prolong_add", lambda: tr.prolong_add(ec, xf))):
            for _ in range(2):
                fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 10
            e0.record()
            for _ in range(reps):
                fn()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            nbytes = 8 * ((N + p) ** 3 * (1 if name == "restrict" else 2) + (N // 2 + p) ** 3)
            print("n=%4d %-12s %-8s %8.4f ms  %6.0f GB/s (algorithmic)" % (
                N + p, name, {False: "per-axis", True: "fused", "v2": "fused-v2"}[fused], ms,
                nbytes / ms * 1e-6), flush=True)
        for op in ("restrict", "prolong"):
            assert tr._want_fused(Vf.npts, op) == {False: None, True: "v1", "v2": "v2"}[fused]
    for k in (True, "v2"):
        dr = (res[k][0] - res[False][0]).abs().max().item() / res[False][0].abs().max().item()
        dp = (res[k][1] - res[False][1]).abs().max().item() / res[False][1].abs().max().item()
        print("   max rel diff %s vs per-axis: restrict %.1e prolong %.1e" % (
            "fused" if k is True else k, dr, dp))
    N //= 2
