"""ORACLE / TEST INFRASTRUCTURE ONLY.  One-rank stand-in for `spl.ddm.cart.Cart`
(call sites: sources/tests/test_kron_dot.py:51-52, sources/kron_product.py:140-141)."""
from mpi4py import MPI


class Cart:
    def __init__(self, npts, pads, periods, reorder=False, comm=None):
        self.npts = tuple(int(n) for n in npts)
        self.pads = tuple(int(p) for p in pads)
        self.periods = tuple(bool(b) for b in periods)
        self.ndim = len(self.npts)
        self.comm_cart = comm if comm is not None else MPI.COMM_WORLD
        self._rank = 0
        self._size = 1
        self.nprocs = [1] * self.ndim
        self.coords = [0] * self.ndim
        self.starts = tuple(0 for _ in self.npts)
        self.ends = tuple(n - 1 for n in self.npts)
        self.global_starts = [[0] for _ in self.npts]
        self.global_ends = [[n - 1] for n in self.npts]
        self.subcomm = [MPI.Comm() for _ in self.npts]
