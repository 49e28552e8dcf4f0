"""Drop-in for /root/reference/sources/kron_product.py (+ the raw-array variants of
/root/reference/pyccel/kron_product.py): Kronecker mat-vec and Kronecker solves on the GPU.

Same names, argument order and return values as the reference.  Dimension-generic
extensions (`kron_dot`, `kron_solve_bnd`) take lists of factors.
"""
import time

import numpy as np
import torch

from . import _lib
from . import bsplines as bs
from . import profiling
from .stencil import (StencilVector, StencilMatrix, KronSumMatrix, DeviceContext, _stream)

__all__ = ["kron_dot_v1", "kron_dot_v2", "kron_dot", "kron_solve_serial", "kron_solve_par",
           "kron_solve_bnd_par", "kron_solve_bnd", "kron_solve_par_bnd_2d",
           "kron_solve_par_bnd_3d", "to_bnd", "BandLU"]


def _band_of(A):
    return A._data if isinstance(A, StencilMatrix) else np.asarray(A, dtype=np.float64)


def kron_dot(As, X):
    """Y = (A_1 (x) ... (x) A_d) X for 1-D banded factors (d = 2, 3)."""
    op = KronSumMatrix([_band_of(A) for A in As])
    return op.dot(X)


def kron_dot_v2(A, B, X):
    """Y = (A kron B) X = A X B^T (/root/reference/sources/kron_product.py:56-89)."""
    return kron_dot([A, B], X)


def kron_dot_v1(A, B, X):
    """Same result as kron_dot_v2 (/root/reference/sources/kron_product.py:10-52 computes it
    entry by entry with ghost exchanges inside the loops)."""
    return kron_dot([A, B], X)


def to_bnd(A):
    """1-D stencil matrix (or dense array) -> LAPACK general band storage (A_bnd, la, ua) with
    A_bnd[la+ua+i-j, j] = A[i, j] (/root/reference/sources/kron_product.py:175-187, which
    raises NameError there; working copy sources/tests/test_kron_solve_bnd.py:30-42)."""
    D = A.toarray() if hasattr(A, "toarray") else np.asarray(A, dtype=float)
    i, j = np.nonzero(D)
    la = int(max(0, (i - j).max()))
    ua = int(max(0, (j - i).max()))
    A_bnd = np.zeros((1 + ua + 2 * la, D.shape[1]))
    A_bnd[la + ua + i - j, j] = D[i, j]
    return A_bnd, la, ua


class BandLU:
    """A dgbtrf factorisation resident on the device."""

    def __init__(self, lub, kl, ku, piv, device):
        self.n = lub.shape[1]
        self.kl, self.ku = int(kl), int(ku)
        if not (0 <= self.kl <= 5 and 0 <= self.ku <= 5):
            raise ValueError("band solve supports kl, ku <= 5")
        assert lub.shape[0] == 2 * self.kl + self.ku + 1
        self._ab_host = np.ascontiguousarray(lub, dtype=np.float64)
        self.ab = torch.as_tensor(self._ab_host, device=device)
        self.piv = torch.as_tensor(np.ascontiguousarray(piv, dtype=np.int32), device=device)
        # dgbtrf made no row interchange (SPD / diagonally dominant bands): the streaming
        # no-pivot kernels apply
        self.nopiv = bool(np.array_equal(np.asarray(piv), np.arange(self.n)))
        self.band = None

    @property
    def piv_ptr(self):
        return None if self.nopiv else self.piv.data_ptr()

    # ---- chunked substitution for few-line problems (2-D grids) ---------------------------------
    def _emulate_chunked(self, y, chunk, warm):
        """Host emulation of band_chunk_kernel (one right-hand side), vectorised over chunks."""
        ab = self._ab_host
        n, kl, ku = self.n, self.kl, self.ku
        kd = kl + ku
        nch = (n + chunk - 1) // chunk
        jb = np.arange(nch) * chunk
        je = np.minimum(n, jb + chunk)
        z = np.zeros((max(kl, 1), nch))
        out = np.zeros(n)
        j0 = np.maximum(0, jb - warm)
        for step in range(chunk + warm):
            j = j0 + step
            act = j < je
            jj = np.where(act, j, 0)
            sv = y[jj].copy()
            for m in range(kl, 0, -1):
                ok = act & (j - m >= j0)
                sv -= np.where(ok, ab[kd + m, np.maximum(jj - m, 0)] * z[m - 1], 0.0)
            st = act & (j >= jb)
            out[jj[st]] = sv[st]
            for m in range(kl - 1, 0, -1):
                z[m] = np.where(act, z[m - 1], z[m])
            if kl > 0:
                z[0] = np.where(act, sv, z[0])
        fwd = out
        out = np.zeros(n)
        w = np.zeros((max(ku, 1), nch))
        jt = np.minimum(n, je + warm)
        for step in range(chunk + warm):
            j = jt - 1 - step
            act = j >= jb
            jj = np.where(act, j, 0)
            sv = fwd[jj].copy()
            for m in range(ku, 0, -1):
                ok = act & (j + m < jt)
                sv -= np.where(ok, ab[kd - m, np.minimum(jj + m, n - 1)] * w[m - 1], 0.0)
            sv = sv * (1.0 / ab[kd, jj])
            st = act & (j < je)
            out[jj[st]] = sv[st]
            for m in range(ku - 1, 0, -1):
                w[m] = np.where(act, w[m - 1], w[m])
            if ku > 0:
                w[0] = np.where(act, sv, w[0])
        return out

    def chunk_plan(self, lines):
        """(chunk, warm) for the chunked kernels, or None when the factor does not allow it.  The
        warm-up length is found by emulating the chunked sweeps on the host and comparing with the
        sequential dgbtrs (<= 1e-14 relative), i.e. the geometric decay is verified, not assumed."""
        if not self.nopiv or self.n < 256:
            return None
        target = max(1, int(round(98304 / max(lines, 1))))
        chunk = max(32, -(-self.n // target))
        chunk = min(chunk, self.n)
        key = chunk
        cache = self.__dict__.setdefault("_chunk_cache", {})
        if key in cache:
            return cache[key]
        from scipy.linalg.lapack import dgbtrs
        rng = np.random.default_rng(12345)
        y = rng.standard_normal(self.n)
        ref, info = dgbtrs(self._ab_host, self.kl, self.ku, y, np.arange(self.n, dtype=np.int32))
        plan = None
        for warm in (32, 48, 64, 96, 128, 192, 256):
            if warm > 4 * chunk:
                break
            x = self._emulate_chunked(y, chunk, warm)
            if np.abs(x - ref).max() <= 1e-14 * np.abs(ref).max():
                plan = (chunk, warm)
                break
        cache[key] = plan
        return plan

    @classmethod
    def from_band(cls, band, device):
        """dgbtrf of an (n, 2p+1) band, trimmed to its true bandwidth first."""
        band = np.asarray(band, dtype=np.float64)
        p = (band.shape[1] - 1) // 2
        q = p
        while q > 0 and not band[:, p - q].any() and not band[:, p + q].any():
            q -= 1
        from . import setup_device as sd
        trimmed = band[:, p - q:p + q + 1]
        if sd.enabled():
            # device LU without row interchanges (SPD / diagonally dominant bands: mass, GLT); a zero
            # pivot falls back to LAPACK's pivoting factorisation on the host
            try:
                lu = cls(*sd.band_lu(trimmed), device)
            except np.linalg.LinAlgError:
                lu = cls(*bs.band_lu(trimmed), device)
        else:
            lu = cls(*bs.band_lu(trimmed), device)
        lu.band = band          # kept for the slab-partitioned (SPIKE) solve, dist.py
        return lu


def _solve_axis(lu, src, dst, axis):
    V = src.space
    shape = tuple(V.local_shape)
    ld = V.ld
    nd = len(shape)
    n = shape[axis]
    assert n == lu.n, "factor size does not match the axis"
    if axis == nd - 1:            # contiguous axis: one line per row of the pitched array
        n_outer, s_outer, s_axis, n_inner = int(np.prod(shape[:-1])), ld, 1, 1
    elif axis == 0:               # slowest axis: lines are (rest of the array) apart
        rest = int(np.prod(shape[1:-1])) * ld if nd > 1 else 1
        n_outer, s_outer, s_axis, n_inner = 1, n * rest, rest, rest
    else:                         # middle axis of a 3-D array
        n_outer, s_outer, s_axis, n_inner = shape[0], shape[1] * ld, ld, ld
    lines = n_outer * (n_inner if axis != nd - 1 else 1)
    plan = lu.chunk_plan(lines) if lines <= CHUNKED_MAX_LINES else None
    if plan is not None:
        # few lines (2-D grids): chunked substitution with verified warm-up, y -> work -> x
        work = torch.empty_like(src.flat)
        _lib.check(_lib.lib().poms_band_solve_axis_chunked(
            src.ptr, dst.ptr, work.data_ptr(), lu.ab.data_ptr(), n, lu.kl, lu.ku, n_outer,
            s_outer, s_axis, n_inner, plan[0], plan[1], plan[1], _stream()),
            "poms_band_solve_axis_chunked")
        return
    _lib.check(_lib.lib().poms_band_solve_axis(
        src.ptr, dst.ptr, lu.ab.data_ptr(), lu.piv_ptr, n, lu.kl, lu.ku, n_outer,
        s_outer, s_axis, n_inner, _stream()), "poms_band_solve_axis")


CHUNKED_MAX_LINES = 32768


def _solve_last_axis_fused(lu, src, work, out, scale, add):
    """Last (contiguous) axis solve with out = [add +] scale * solution."""
    V = src.space
    shape = tuple(V.local_shape)
    n = shape[-1]
    assert n == lu.n and lu.nopiv
    _lib.check(_lib.lib().poms_band_solve_axis_fused(
        src.ptr, work.ptr, lu.ab.data_ptr(), n, lu.kl, lu.ku, int(np.prod(shape[:-1])), V.ld,
        float(scale), add.ptr if add is not None else None, out.ptr, _stream()),
        "poms_band_solve_axis_fused")


def kron_solve_bnd_update(factors, Y, work, out, scale, add=None):
    """out = [add +] scale * (A_1 (x) .. (x) A_d)^-1 Y with the scaling / accumulation fused into
    the last line solve (EXTENSION: smoother update of mg.Hierarchy.smooth).  `work` receives
    intermediates; factors must be no-pivot BandLU objects."""
    V = Y.space
    lus = list(factors)
    if V.slab is not None and V.slab.size > 1:
        from .dist import kron_solve_bnd_slab
        kron_solve_bnd_slab(lus[:1] + [None] * (len(lus) - 1), Y, work, only_axis1=True)
        src = work
        first = 1
    else:
        src = Y
        first = 0
    for ax in range(first, len(lus) - 1):
        with profiling.region("band_solve_axis%d" % (ax + 1), 16 * V.local_size):
            _solve_axis(lus[ax], src, work, ax)
        src = work
    last_lines = int(np.prod(V.local_shape[:-1]))
    if last_lines <= CHUNKED_MAX_LINES and lus[-1].chunk_plan(last_lines) is not None:
        # few lines: chunked solve of the last axis, then the update as one BLAS-1 pass
        with profiling.region("band_solve_axis%d" % len(lus), 16 * V.local_size):
            _solve_axis(lus[-1], src, work, len(lus) - 1)
        with profiling.region("smoother_update", 24 * V.local_size):
            L = _lib.lib()
            if add is None:
                _lib.check(L.poms_axpby(out.ptr, float(scale), work.ptr, 0.0, None, out.n_owned,
                                        _stream()), "poms_axpby")
            else:
                _lib.check(L.poms_axpby(out.ptr, 1.0, add.ptr, float(scale), work.ptr,
                                        out.n_owned, _stream()), "poms_axpby")
        return out
    with profiling.region("band_solve_axis%d" % len(lus), (16 if add is None else 24) * V.local_size):
        _solve_last_axis_fused(lus[-1], src, work, out, scale, add)
    return out


def kron_solve_bnd(factors, Y, X=None):
    """X = (A_1 (x) .. (x) A_d)^-1 Y with pre-factored banded A_a (BandLU or the reference's
    [A_bnd, la, ua, piv] lists); dgbtrs sweeps along axis 1, then 2 (, then 3) as in
    /root/reference/pyccel/pyccel_functions.py:226-244."""
    V = Y.space
    if V.slab is not None and V.slab.size > 1:
        from .dist import kron_solve_bnd_slab
        return kron_solve_bnd_slab(factors, Y, X)
    lus = [f if isinstance(f, BandLU) else BandLU(f[0], f[1], f[2], f[3], V.device)
           for f in factors]
    if X is None:
        X = StencilVector(V, zero=False)
    src = Y
    for ax, lu in enumerate(lus):
        # algorithmic bytes of one dgbtrs sweep pair: forward (read y, write t) + backward
        # (read t, write x) = 32 B/DOF unfused; 16 B/DOF is the fused bound
        with profiling.region("band_solve_axis%d" % (ax + 1), 16 * V.local_size):
            _solve_axis(lu, src, X, ax)
        src = X
    return X


def kron_solve_bnd_par(A, B, Y):
    """(X, elapsed) = solve (A kron B) X = Y with A = [A_bnd, la, ua, A_piv] as returned by
    dgbtrf (/root/reference/sources/kron_product.py:191-239)."""
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    X = kron_solve_bnd([A, B], Y)
    torch.cuda.synchronize()
    return X, time.perf_counter() - t0


def _cached_lu(A, device):
    """dgbtrf of a 1-D StencilMatrix, cached on the object until its entries change."""
    if not isinstance(A, StencilMatrix):
        return BandLU.from_band(np.asarray(A, dtype=np.float64), device)
    cache = A.__dict__.setdefault("_lu_cache", {})
    key = str(device)
    ent = cache.get(key)
    if ent is None or not np.array_equal(ent[0], A._data):
        cache[key] = (A._data.copy(), BandLU.from_band(A._data, device))
    return cache[key][1]


def kron_solve_serial(A, B, Y):
    """X = (A kron B)^-1 Y (/root/reference/sources/kron_product.py:93-117).  The reference
    densifies and dgetrf-factors both matrices on every call; here the banded factors are
    computed once per matrix (dgbtrf) and the line solves run on the device."""
    dev = Y.space.device
    return kron_solve_bnd([_cached_lu(A, dev), _cached_lu(B, dev)], Y)


def kron_solve_par(A, B, Y):
    """MPI variant of kron_solve_serial (/root/reference/sources/kron_product.py:121-170): same
    result; under a slab partition the axis-1 lines are solved after a slab<->pencil exchange
    instead of one Allgatherv per line."""
    return kron_solve_serial(A, B, Y)


def kron_solve_par_bnd_2d(A_bnd, la, ua, B_bnd, lb, ub, Y, X, with_pycc=False):
    """/root/reference/pyccel/kron_product.py:135-169: UNFACTORED LAPACK bands in, result
    written into X (the pyccel kernel re-factorises on every call,
    pyccel/pyccel_functions.py:146-147)."""
    from scipy.linalg.lapack import dgbtrf
    fa = dgbtrf(np.array(A_bnd, dtype=float), la, ua)
    fb = dgbtrf(np.array(B_bnd, dtype=float), lb, ub)
    return kron_solve_bnd([[fa[0], la, ua, fa[1]], [fb[0], lb, ub, fb[1]]], Y, X)


def kron_solve_par_bnd_3d(A_bnd, la, ua, B_bnd, lb, ub, C_bnd, lc, uc, Y, X):
    """/root/reference/pyccel/kron_product.py:171-192 (the only 3-D code of the reference)."""
    from scipy.linalg.lapack import dgbtrf
    fs = []
    for ab, l, u in ((A_bnd, la, ua), (B_bnd, lb, ub), (C_bnd, lc, uc)):
        f = dgbtrf(np.array(ab, dtype=float), l, u)
        fs.append([f[0], l, u, f[1]])
    return kron_solve_bnd(fs, Y, X)
