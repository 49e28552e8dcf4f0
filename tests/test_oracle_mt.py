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
