"""A/B timing of the 3-D TMA mat-vec per level of the C5 hierarchy: operator (resid epilogue), the two
smoother factors S1 (store) and S2 (axpy), each with the automatic axis-1 chunk and a sweep of fixed
chunks, checked against the generic kernel -- run on the GPU box:
    python tests/gpu_ab_chunk.py [N] [p] [chunks,comma,separated]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from poms_b200 import _lib
from poms_b200.mg import Hierarchy
from poms_b200.stencil import StencilVector, DeviceContext, EPI_STORE, EPI_RESID, EPI_AXPY

N = int(sys.argv[1]) if len(sys.argv) > 1 else 512
p = int(sys.argv[2]) if len(sys.argv) > 2 else 3
chunks = [int(c) for c in sys.argv[3].split(",")] if len(sys.argv) > 3 else [0]
dev = torch.device("cuda", 0)
h = Hierarchy(p, N, ndim=3, Nc=32, device=dev, smoother="glt_poly")
ctx = DeviceContext.get(dev)
L = _lib.lib()
g = torch.Generator(device=dev).manual_seed(0)


def timed(fn, reps):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for lv in h.levels[:-1]:
    V = lv.V
    x, y, yg, b = (StencilVector(V) for _ in range(4))
    x.data.copy_(torch.randn(V.npts, generator=g, dtype=torch.float64, device=dev))
    b.data.copy_(torch.randn(V.npts, generator=g, dtype=torch.float64, device=dev))
    dof = V.local_size
    reps = 10 if dof > 5e7 else 30
    for name, M, epi, nbytes in (("A/resid", lv.A, EPI_RESID, 24), ("S1/store", lv.S1, EPI_STORE, 16),
                                 ("S2/axpy", lv.S2, EPI_AXPY, 24)):
        L.poms_set_force_generic(1)
        M.apply(x, yg, epi, b=b, omega=0.5, dot_ptr=ctx.sptr(20))
        L.poms_set_force_generic(0)
        line = "n=%4d %-9s" % (V.npts[0], name)
        for c in chunks:
            L.poms_set_matvec3d_chunk(c)
            ms = timed(lambda: M.apply(x, y, epi, b=b, omega=0.5, dot_ptr=ctx.sptr(21)), reps)
            diff = (y.data - yg.data).abs().max().item() / yg.data.abs().max().item()
            dd = abs(ctx.scal[21].item() - ctx.scal[20].item()) / abs(ctx.scal[20].item())
            line += "  c=%-3d %7.4f ms %6.0f GB/s%s" % (c, ms, nbytes * dof / ms / 1e6,
                                                       "" if diff < 1e-13 and dd < 1e-12 else " BAD(%.1e,%.1e)" % (diff, dd))
        print(line, flush=True)
L.poms_set_matvec3d_chunk(0)
