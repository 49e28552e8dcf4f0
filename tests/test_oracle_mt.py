"""The threaded CPU port used as the timed baseline computes what the pinned oracle computes."""
import numpy as np
import pytest

from oracle import poms_oracle as po

mt = pytest.importorskip("oracle.poms_oracle_mt")


@pytest.mark.parametrize("smoother", ["glt", "glt_poly"])
@pytest.mark.parametrize("p,N", [(3, (24, 16)), (2, (8, 16, 12))])
def test_threaded_port_matches_oracle(p, N, smoother):
    h0 = po.MGHierarchy(p, list(N), smoother=smoother)
    h1 = mt.MGHierarchyMT(p, list(N), smoother=smoother)
    rng = np.random.default_rng(0)
    X = rng.standard_normal(h0.levels[0]["A"].npts)
    assert np.abs(h1.levels[0]["A"].dot(X) - h0.levels[0]["A"].dot(X)).max() < 1e-12 * np.abs(X).max() * 1e3
    if smoother == "glt":
        z0 = po.kron_solve_banded(h0.levels[0]["glt"], X)
        z1 = mt.kron_solve_banded(h1.levels[0]["glt"], X)
        assert np.abs(z0 - z1).max() < 1e-12 * np.abs(z0).max()
    b = np.ones(X.shape)
    x0, i0 = h0.mg_pcg(b)
    x1, i1 = h1.mg_pcg(b)
    assert i0["niter"] == i1["niter"]
    assert np.allclose(i0["history"], i1["history"], rtol=1e-6)
    assert np.abs(x0 - x1).max() < 1e-9 * np.abs(x0).max()


@pytest.mark.parametrize("p,N,lengths", [(3, (16, 16, 16), None), (2, (32, 8, 8), (4.0, 1.0, 1.0)), (3, (32, 16), None)])
def test_coarse_solve_by_eigenpairs_equals_the_dense_inverse(p, N, lengths, monkeypatch):
    """Above MGHierarchy.DENSE_COARSE_MAX unknowns the oracle's exact coarse solve goes through the
    1-D generalised eigenpairs instead of a dense inverse (the C3 parity test stops at 35^3): same
    solve, same MG-PCG iteration counts."""
    kw = {} if lengths is None else {"lengths": list(lengths)}
    h_dense = po.MGHierarchy(p, list(N), **kw)
    assert h_dense._fd is None
    monkeypatch.setattr(po.MGHierarchy, "DENSE_COARSE_MAX", 0)
    h_fd = mt.MGHierarchyMT(p, list(N), **kw)
    assert h_fd._fd is not None and h_fd.Ainv_c is None
    b = np.random.default_rng(0).standard_normal(h_dense.levels[-1]["A"].npts)
    xd, xf = h_dense.coarse_solve(b), h_fd.coarse_solve(b)
    assert np.abs(xd - xf).max() < 1e-12 * np.abs(xd).max()
    assert np.abs(h_dense.levels[-1]["A"].dot(xf) - b).max() < 1e-11 * np.abs(b).max()
    bb = np.ones(h_dense.levels[0]["A"].npts)
    x0, i0 = h_dense.mg_pcg(bb)
    x1, i1 = h_fd.mg_pcg(bb)
    assert i0["niter"] == i1["niter"] and i0["restarts"] == i1["restarts"]
    assert np.abs(x0 - x1).max() < 1e-10 * np.abs(x0).max()
