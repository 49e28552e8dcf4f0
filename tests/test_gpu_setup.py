"""Device-side 1-D setup (csrc/poms_setup.cu, poms_b200/setup_device.py) against the host NumPy / SciPy
routines the golden-vector tests pin: band assembly, knot-insertion rows, banded LU, generalised
eigenproblems -- and a whole hierarchy built on the device solving to the same iteration count."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda", 0)


def _knots(p, N, uniform):
    from poms_b200 import bsplines as bs
    T = bs.make_open_knots(p, N + p)
    if not uniform:                      # graded interior breakpoints, still an open knot vector
        inner = np.linspace(0.0, 1.0, N + 1)[1:-1] ** 1.7
        T = np.concatenate([np.zeros(p + 1), inner, np.ones(p + 1)])
    return T


@pytest.mark.parametrize("p", [1, 2, 3, 4, 5])
@pytest.mark.parametrize("N,uniform", [(8, True), (37, True), (23, False), (512, True)])
def test_assembly_matches_host(dev, p, N, uniform):
    from poms_b200 import bsplines as bs, setup_device as sd
    T = _knots(p, N, uniform)
    Mh, Kh = bs.assemble_1d_bands(p, T)
    with sd.device_setup(True, dev):
        Md, Kd = sd.assemble_1d_bands(p, T)
    assert np.abs(Md - Mh).max() <= 1e-13 * np.abs(Mh).max()
    assert np.abs(Kd - Kh).max() <= 1e-13 * np.abs(Kh).max()
    if uniform and N + p > 4 * p + 1:    # Toeplitz interior rows are bit-identical, like the host's
        n = N + p
        assert np.all(Md[2 * p:n - 2 * p] == Md[n // 2]) and np.all(Kd[2 * p:n - 2 * p] == Kd[n // 2])


@pytest.mark.parametrize("p", [1, 2, 3, 5])
@pytest.mark.parametrize("Nc,ratio", [(4, 2), (8, 2), (5, 4), (16, 8)])
def test_knot_insertion_rows_match_host(dev, p, Nc, ratio):
    from poms_b200 import bsplines as bs, setup_device as sd
    Tc, Tf = bs.make_open_knots(p, Nc + p), bs.make_open_knots(p, Nc * ratio + p)
    sh, ch, nc = bs.knot_insertion_rows(Tc, Tf, p)
    with sd.device_setup(True, dev):
        sdv, cd, ncd = sd.knot_insertion_rows(Tc, Tf, p)
    assert ncd == nc and np.array_equal(sdv, sh)
    assert np.abs(cd - ch).max() <= 1e-15
    assert np.abs(cd.sum(axis=1) - 1.0).max() < 1e-14        # partition of unity of discrete B-splines


@pytest.mark.parametrize("p,n", [(2, 40), (3, 67), (5, 133), (3, 515)])
def test_band_lu_matches_lapack(dev, p, n):
    from poms_b200 import bsplines as bs, setup_device as sd
    band = bs.glt_band(p, n, degree=max(2 * p - 1, 1))
    q = (band.shape[1] - 1) // 2
    while q > 0 and not band[:, (band.shape[1] - 1) // 2 - q].any():
        q -= 1
    c = (band.shape[1] - 1) // 2
    band = band[:, c - q:c + q + 1]
    lh, kl, ku, piv = bs.band_lu(band)
    assert np.array_equal(piv, np.arange(n))                   # no interchanges: comparable factors
    with sd.device_setup(True, dev):
        ld_, kld, kud, pivd = sd.band_lu(band)
    assert (kld, kud) == (kl, ku)
    assert np.abs(ld_ - lh).max() <= 1e-13 * np.abs(lh).max()


def test_generalised_eigenproblem_on_device(dev):
    from scipy.linalg import eigh
    from poms_b200 import bsplines as bs, setup_device as sd
    M, K = bs.assemble_1d_bands(3, bs.make_open_knots(3, 35))
    Md, Kd = bs.band_to_dense(M), bs.band_to_dense(K + M)
    w, Q = eigh(Kd, Md)
    with sd.device_setup(True, dev):
        wd, Qd = sd.gen_eigh(Kd, Md)
        lmax = sd.gen_eig_max(K + M, M)
    assert np.abs(wd - w).max() <= 1e-10 * w.max()
    assert np.abs(Qd.T @ Md @ Qd - np.eye(35)).max() < 1e-10
    assert np.abs(Kd @ Qd - Md @ Qd * wd).max() < 1e-8 * w.max()
    assert abs(lmax - w[-1]) <= 1e-10 * w[-1]


@pytest.mark.parametrize("p,N,smoother", [(3, (32, 32), "glt"), (3, (16, 16, 16), "glt_poly"), (2, (64, 16), "glt")])
def test_hierarchy_built_on_device_solves_like_host(dev, p, N, smoother):
    from poms_b200.mg import Hierarchy, mg_pcg
    from poms_b200.stencil import StencilVector
    out = {}
    for mode in ("host", "device"):
        h = Hierarchy(p, list(N), device=dev, smoother=smoother, nu=1, setup=mode)
        assert h.setup == mode
        b = StencilVector(h.levels[0].V)
        b.data.fill_(1.0)
        x, info = mg_pcg(h, b, tol=1e-10, maxiter=100)
        out[mode] = (x.toarray().copy(), info, [lv.lmax for lv in h.levels[:-1]])
    assert out["device"][1]["niter"] == out["host"][1]["niter"]
    assert np.allclose(out["device"][2], out["host"][2], rtol=1e-9)
    assert np.abs(out["device"][0] - out["host"][0]).max() <= 1e-9 * np.abs(out["host"][0]).max()
