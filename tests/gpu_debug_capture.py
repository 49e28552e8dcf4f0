"""Diagnostic (not a pytest test): locate the call that invalidates CUDA-graph capture of the PCG
iteration.  Usage: python tests/gpu_debug_capture.py [order]   order = 'p1first' | 'p3first'."""
import os
import sys
import traceback

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from poms_b200 import _lib  # noqa: E402
from poms_b200.mg import Hierarchy, mg_pcg  # noqa: E402
from poms_b200.stencil import StencilVector, DeviceContext, EPI_STORE, _stream  # noqa: E402
from poms_b200 import solvers  # noqa: E402

dev = torch.device("cuda", 0)
cases = [(3, (16, 16, 16), "glt"), (1, (16, 16, 16), "glt")]
if len(sys.argv) > 1 and sys.argv[1] == "p1first":
    cases = cases[::-1]


def piecewise(h):
    lv = h.levels[0]
    ctx = DeviceContext.get(h.device)
    L = _lib.lib()
    x, r, p, q = (lv.ws(n) for n in ("dbg_x", "dbg_r", "dbg_p", "dbg_q"))
    p.flat.fill_(1.0)

    def f_apply():
        lv.A.apply(p, q, EPI_STORE, dot_ptr=ctx.sptr(solvers.S_PQ))

    def f_apply_nodot():
        lv.A.apply(p, q, EPI_STORE)

    def f_cg():
        _lib.check(L.poms_cg_update(x.ptr, r.ptr, p.ptr, q.ptr, x.n_owned, ctx.sptr(solvers.S_SR0),
                                    ctx.sptr(solvers.S_PQ), ctx.sptr(solvers.S_RR), ctx.ws_ptr,
                                    _stream()), "poms_cg_update")

    for name, f in (("apply+dot", f_apply), ("apply", f_apply_nodot), ("cg_update", f_cg)):
        f()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        try:
            with torch.cuda.graph(g):
                f()
            g.replay()
            torch.cuda.synchronize()
            print("   capture of %-10s ok" % name, flush=True)
        except Exception as exc:
            print("   capture of %-10s FAILED: %s" % (name, str(exc).splitlines()[0]), flush=True)
            try:
                torch.cuda.synchronize()
            except Exception:
                pass


for p, N, sm in cases:
    print("case p=%d N=%s %s" % (p, N, sm), flush=True)
    h = Hierarchy(p, list(N), device=dev, smoother=sm, nu=1)
    b = StencilVector(h.levels[0].V)
    b.data.fill_(1.0)
    piecewise(h)
    try:
        x, info = mg_pcg(h, b, tol=1e-10, maxiter=100)
        print("   mg_pcg ok: niter %d graphed %s" % (info["niter"], info.get("graphed")), flush=True)
    except Exception:
        traceback.print_exc()
        print("   mg_pcg FAILED", flush=True)
