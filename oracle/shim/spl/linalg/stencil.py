"""ORACLE / TEST INFRASTRUCTURE ONLY -- never imported by the product (poms_b200/).

Serial NumPy stand-in for `spl.linalg.stencil` (third-party, un-vendored, unpinned in
/root/reference/requirements.txt:4).  It provides the duck-typed surface the reference's
hot-path modules use with the semantics fixed by the reference's
own call sites:

* vectors are indexed with GLOBAL indices including the p-wide ghost range
  (sources/kron_product.py:82 reads `X[j1, i2-p2:i2+p2+1]` with j1 in the ghost rows);
  storage `_data` has shape npts_local + 2*pads and the interior is `_data[p:-p]`
  (sources/utils.py:97-99);
* 1-D matrices are indexed `[i, k]`, k in [-p, p] the diagonal offset
  (sources/utils.py:7-16; pyccel/pyccel_functions.py:15,19 use column k+p);
* `M.dot(v)[i] = sum_k M[i,k] * v[i+k]` (slides/content.tex:285-290);
* `v.dot(w)` sums over owned entries only; `+ - *` act on whole storage.

One rank, non-periodic only: `update_ghost_regions` is a no-op.
"""
import numpy as np
from scipy.sparse import coo_matrix


class StencilVectorSpace:
    def __init__(self, *args, **kwargs):
        from spl.ddm.cart import Cart

        if len(args) == 1 and isinstance(args[0], Cart):
            cart = args[0]
        else:
            npts, pads, periods = (list(args) + [None] * 3)[:3]
            npts = kwargs.get("npts", npts)
            pads = kwargs.get("pads", pads)
            periods = kwargs.get("periods", periods) or [False] * len(npts)
            cart = Cart(npts=npts, pads=pads, periods=periods, reorder=False, comm=None)
        assert not any(cart.periods), "shim: periodic spaces not supported"
        self._cart = cart
        self.npts = tuple(cart.npts)
        self.pads = tuple(cart.pads)
        self.periods = tuple(cart.periods)
        self.starts = tuple(cart.starts)
        self.ends = tuple(cart.ends)
        self.ndim = len(self.npts)
        self._mpi_type = None

    @property
    def cart(self):
        return self._cart

    @property
    def dimension(self):
        return int(np.prod(self.npts))

    def zeros(self):
        return StencilVector(self)


def _shift(key, starts, pads):
    if not isinstance(key, tuple):
        key = (key,)
    out = []
    for i, s, p in zip(key, starts, pads):
        if isinstance(i, slice):
            a = None if i.start is None else i.start - s + p
            b = None if i.stop is None else i.stop - s + p
            out.append(slice(a, b, i.step))
        else:
            out.append(i - s + p)
    return tuple(out)


class StencilVector:
    def __init__(self, V):
        self._space = V
        self._data = np.zeros([e - s + 1 + 2 * p for s, e, p in zip(V.starts, V.ends, V.pads)])

    @property
    def space(self):
        return self._space

    @property
    def starts(self):
        return self._space.starts

    @property
    def ends(self):
        return self._space.ends

    @property
    def pads(self):
        return self._space.pads

    @property
    def shape(self):
        return (self._space.dimension,)

    def _interior(self):
        return tuple(slice(p, -p) for p in self._space.pads)

    def __getitem__(self, key):
        return self._data[_shift(key, self.starts, self.pads)]

    def __setitem__(self, key, value):
        self._data[_shift(key, self.starts, self.pads)] = value

    def copy(self):
        w = StencilVector(self._space)
        w._data[...] = self._data
        return w

    def toarray(self):
        return self._data[self._interior()].flatten()

    def dot(self, other):
        idx = self._interior()
        return float(np.dot(self._data[idx].flatten(), other._data[idx].flatten()))

    def update_ghost_regions(self, direction=None):
        return None

    def _new(self, data):
        w = StencilVector(self._space)
        w._data[...] = data
        return w

    def __mul__(self, a):
        return self._new(self._data * a)

    __rmul__ = __mul__

    def __add__(self, v):
        return self._new(self._data + v._data)

    def __sub__(self, v):
        return self._new(self._data - v._data)

    def __neg__(self):
        return self._new(-self._data)


class StencilMatrix:
    """Rows are owned entries of the codomain; `_data[i - s, ..., k + p, ...]`."""

    def __init__(self, V, W):
        assert V.npts == W.npts and V.pads == W.pads
        self._domain = V
        self._codomain = W
        self.ndim = V.ndim
        self.starts = V.starts
        self.ends = V.ends
        self.pads = V.pads
        dims = [e - s + 1 for s, e in zip(V.starts, V.ends)]
        diags = [2 * p + 1 for p in V.pads]
        self._data = np.zeros(dims + diags)

    @property
    def domain(self):
        return self._domain

    @property
    def codomain(self):
        return self._codomain

    @property
    def shape(self):
        n = self._domain.dimension
        return (n, n)

    def _index(self, key):
        assert isinstance(key, tuple) and len(key) == 2 * self.ndim
        rows = key[: self.ndim]
        offs = key[self.ndim:]
        out = []
        for i, s in zip(rows, self.starts):
            if isinstance(i, slice):
                a = None if i.start is None else i.start - s
                b = None if i.stop is None else i.stop - s
                out.append(slice(a, b, i.step))
            else:
                out.append(i - s)
        for k, p in zip(offs, self.pads):
            if isinstance(k, slice):
                a = None if k.start is None else k.start + p
                b = None if k.stop is None else k.stop + p
                out.append(slice(a, b, k.step))
            else:
                out.append(k + p)
        return tuple(out)

    def __getitem__(self, key):
        return self._data[self._index(key)]

    def __setitem__(self, key, value):
        self._data[self._index(key)] = value

    def remove_spurious_entries(self):
        """Zero the entries whose column i+k falls outside [0, n) (non-periodic)."""
        for d, (n, p, s, e) in enumerate(
            zip(self._domain.npts, self.pads, self.starts, self.ends)
        ):
            i = np.arange(s, e + 1)[:, None]
            k = np.arange(-p, p + 1)[None, :]
            bad = (i + k < 0) | (i + k >= n)
            shape = [1] * (2 * self.ndim)
            shape[d] = bad.shape[0]
            shape[self.ndim + d] = bad.shape[1]
            self._data[np.broadcast_to(bad.reshape(shape), self._data.shape)] = 0.0

    def dot(self, v):
        out = StencilVector(self._codomain)
        nd = self.ndim
        pads = self.pads
        dims = self._data.shape[:nd]
        acc = np.zeros(dims)
        for ks in np.ndindex(*[2 * p + 1 for p in pads]):
            src = tuple(slice(k, k + n) for k, n in zip(ks, dims))
            acc += self._data[(Ellipsis,) + ks] * v._data[src]
        out._data[tuple(slice(p, -p) for p in pads)] = acc
        return out

    def tocoo(self):
        nd = self.ndim
        npts = self._domain.npts
        rows, cols, vals = [], [], []
        grids = np.meshgrid(
            *[np.arange(s, e + 1) for s, e in zip(self.starts, self.ends)], indexing="ij"
        )
        for ks in np.ndindex(*[2 * p + 1 for p in self.pads]):
            offs = [k - p for k, p in zip(ks, self.pads)]
            js = [g + o for g, o in zip(grids, offs)]
            ok = np.ones(grids[0].shape, dtype=bool)
            for j, n in zip(js, npts):
                ok &= (j >= 0) & (j < n)
            v = self._data[(Ellipsis,) + ks]
            ok &= v != 0.0
            rows.append(np.ravel_multi_index([g[ok] for g in grids], npts))
            cols.append(np.ravel_multi_index([j[ok] for j in js], npts))
            vals.append(v[ok])
        n = int(np.prod(npts))
        return coo_matrix(
            (np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(n, n)
        )

    def tocsr(self):
        return self.tocoo().tocsr()

    def toarray(self):
        return self.tocoo().toarray()
