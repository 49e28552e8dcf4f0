"""Slab partition of the tensor grid along axis 1 (the slowest axis) across the GPUs of one box.

Replaces the reference's MPI layer for the hot path:
  * `update_ghost_regions` (spl; /root/reference/sources/kron_product.py:76,87, solvers.py:162,215)
        -> `Slab.exchange`: p contiguous planes to each neighbour (NCCL send/recv; the slab axis
           is the slowest one, so a halo is one contiguous block and needs no pack kernel);
  * the allreduce inside `StencilVector.dot` -> `Slab.allreduce_sum` on a device scalar;
  * one `Allgatherv` PER LINE in kron_solve_par / kron_solve_bnd_par
    (/root/reference/sources/kron_product.py:156,224) -> a SPIKE partitioned banded solve: every
    slab solves its own diagonal block, ONE all-gather of the 2q interface planes, a tiny replicated
    reduced system, and a correction that only touches the planes where the (exponentially
    decaying) spikes are above rounding;
  * `comm.allreduce(rc)` of the coarse residual (/root/reference/sources/mg_jac.py:95) ->
    `Slab.allgather_planes` at the level where the hierarchy becomes replicated.

Every method takes plain torch tensors whose first dimension is the plane index, so the
communication logic runs unchanged on CPU tensors with the gloo backend (tests/test_dist_gloo.py).
"""
import os

import numpy as np
import torch
import torch.distributed as dist

from . import profiling


def block_bounds(n, size, rank):
    """Inclusive [s, e] of the `rank`-th of `size` nearly equal contiguous blocks of range(n)."""
    base, rem = divmod(n, size)
    s = rank * base + min(rank, rem)
    e = s + base + (1 if rank < rem else 0) - 1
    return s, e


def p2p_offsets(n_glob, size, rank, w):
    """Plane offsets of a peer-store halo exchange of width `w` (host logic of `Slab._exchange_p2p`,
    pure so that it can be tested without a GPU).  Every rank stores [glo ghost | n_own owned | ghi
    ghost] planes with glo = w except on rank 0 and ghi = w except on the last rank.  Returns the
    plane index (in the rank's own storage) of the w owned planes it pushes to the lower / upper
    neighbour (`src_lo`, `src_hi`) and the plane index IN THE NEIGHBOUR'S storage where they land
    (`dst_lo`: the lower neighbour's upper ghost planes, `dst_hi`: the upper neighbour's lower ghost
    planes); None where there is no neighbour."""
    s, e = block_bounds(n_glob, size, rank)
    n_own = e - s + 1
    glo = w if rank > 0 else 0
    ghi = w if rank < size - 1 else 0
    out = {"n_own": n_own, "glo": glo, "ghi": ghi, "src_lo": None, "dst_lo": None, "src_hi": None,
           "dst_hi": None}
    if rank > 0:
        assert n_own >= w
        s_, e_ = block_bounds(n_glob, size, rank - 1)
        glo_nb = w if rank - 1 > 0 else 0
        out["src_lo"] = glo
        out["dst_lo"] = glo_nb + (e_ - s_ + 1)
    if rank < size - 1:
        assert n_own >= w
        out["src_hi"] = glo + n_own - w
        out["dst_hi"] = 0
    return out


class _DeviceBlob:
    """A raw device allocation exposed through __cuda_array_interface__ so that torch can view it
    (torch.as_tensor) without owning it."""

    def __init__(self, ptr, n_doubles):
        self.__cuda_array_interface__ = {"shape": (int(n_doubles),), "typestr": "<f8",
                                         "data": (int(ptr), False), "version": 2}


class PeerArena:
    """IPC-shared device memory for the vectors whose ghost planes are exchanged: every rank
    allocates the same sequence of equally sized blocks (sizes are computed from the LARGEST slab),
    so a vector sits at the same (chunk, offset) on every rank and a rank can address its
    neighbours' copy through the chunk's mapped base pointer.  Blocks are never freed: the arena
    holds the persistent work vectors of a hierarchy (`Level.ws`).  Chunks are opened lazily
    (collective: handles travel with all_gather_object)."""

    ALIGN = 512

    def __init__(self, slab):
        import ctypes as C
        from . import _lib
        self.slab = slab
        self.L = _lib.lib()
        self.C = C
        self.chunk_bytes = int(float(os.environ.get("POMS_B200_ARENA_CHUNK_GB", "8")) * (1 << 30))
        self.max_bytes = int(float(os.environ.get("POMS_B200_ARENA_MAX_GB", "96")) * (1 << 30))
        self.total = 0
        self.chunks = []          # dicts: base, size, used, tensor, lo (peer base), hi (peer base)
        self.ok = True
        fl = self._new_chunk(4096)
        self.flags = fl           # word 0.. : this rank's handshake flags
        self.ok = fl is not None

    def _new_chunk(self, nbytes):
        C, L, slab = self.C, self.L, self.slab
        ptr = C.c_void_p()
        hb = L.poms_ipc_handle_bytes()
        handle = C.create_string_buffer(hb)
        good = L.poms_ipc_alloc(nbytes, C.byref(ptr)) == 0
        good = good and L.poms_ipc_get_handle(ptr, handle) == 0
        mine = handle.raw if good else None
        allh = [None] * slab.size
        dist.all_gather_object(allh, mine, group=slab.group)
        ch = {"base": ptr.value if good else 0, "size": nbytes, "used": 0, "lo": 0, "hi": 0}
        if all(h is not None for h in allh):
            for side, r in (("lo", slab.rank - 1), ("hi", slab.rank + 1)):
                if 0 <= r < slab.size:
                    q = C.c_void_p()
                    if L.poms_ipc_open(C.c_char_p(allh[r]), C.byref(q)) == 0:
                        ch[side] = q.value
                    else:
                        good = False
        else:
            good = False
        # the decision must be the same on every rank
        flag = torch.tensor([1 if good else 0], device=slab.device, dtype=torch.int32)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=slab.group)
        if int(flag.item()) == 0:
            return None
        ch["tensor"] = torch.as_tensor(_DeviceBlob(ch["base"], nbytes // 8), device=slab.device)
        self.chunks.append(ch)
        self.total += nbytes
        return ch

    def alloc(self, n_doubles):
        """(chunk index, byte offset, 1-D float64 view) or None when the arena is exhausted/unusable.
        COLLECTIVE when a new chunk is needed (same call sequence on every rank)."""
        if not self.ok:
            return None
        nbytes = -(-8 * int(n_doubles) // self.ALIGN) * self.ALIGN
        for i, ch in enumerate(self.chunks[1:], start=1):
            if ch["size"] - ch["used"] >= nbytes:
                break
        else:
            want = max(self.chunk_bytes, nbytes)
            if self.total + want > self.max_bytes:
                return None
            ch = self._new_chunk(want)
            if ch is None:
                self.ok = False
                return None
            i = len(self.chunks) - 1
        off = ch["used"]
        ch["used"] += nbytes
        view = ch["tensor"][off // 8: off // 8 + int(n_doubles)]
        view.zero_()
        return i, off, view


class Slab:
    def __init__(self, group=None, device=None):
        self.group = group if group is not None else dist.group.WORLD
        self.rank = dist.get_rank(self.group)
        self.size = dist.get_world_size(self.group)
        self.device = device
        # POMS_B200_OVERLAP=1: run the halo exchange of a 3-D mat-vec on a second stream while the
        # ghost-free planes are computed.  OFF by default -- measured slower on 2 B200s (C5: 445 ms
        # vs 337 ms): the mat-vec fills every SM's register file (2 x 256 threads x 128 registers), so
        # NCCL's copy kernel only starts when a CTA retires, ~1/8 of the kernel later, which is longer
        # than the 35 us exchange it was meant to hide.
        self.overlap = (device is not None and torch.device(device).type == "cuda"
                        and os.environ.get("POMS_B200_OVERLAP") == "1")
        self._comm_stream = None
        # POMS_B200_P2P=0 keeps every halo exchange on NCCL send/recv.  Default on CUDA: the
        # persistent work vectors live in an IPC-shared arena and ghost planes move by peer stores
        # (csrc/poms_extra.cu); vectors outside the arena still go through NCCL.
        self.arena = None
        self._want_p2p = (device is not None and torch.device(device).type == "cuda"
                          and self.size > 1 and os.environ.get("POMS_B200_P2P", "1") != "0")

    def get_arena(self):
        """The rank's peer arena (created on first use; collective), or None."""
        if self._want_p2p and self.arena is None:
            a = PeerArena(self)
            self._want_p2p = a.ok
            self.arena = a if a.ok else None
        return self.arena

    def alloc_planes(self, n_planes_max, plane_shape):
        """Storage for a vector with at most `n_planes_max` planes (ghosts included; the same number
        on every rank) from the peer arena: (tensor view of that many planes, arena key) or None."""
        a = self.get_arena()
        if a is None:
            return None
        per = int(np.prod(plane_shape))
        got = a.alloc(n_planes_max * per)
        if got is None:
            return None
        i, off, view = got
        return view.view((n_planes_max,) + tuple(plane_shape)), (i, off)

    # ---- partition ---------------------------------------------------------------------------
    def bounds(self, n, rank=None):
        return block_bounds(n, self.size, self.rank if rank is None else rank)

    def table(self, n):
        return [block_bounds(n, self.size, r) for r in range(self.size)]

    # ---- collectives -------------------------------------------------------------------------
    def allreduce_sum(self, t):
        if self.size > 1:
            with profiling.region("allreduce_scalar", 0):
                dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return t

    def _peer(self, r):
        return dist.get_global_rank(self.group, r) if self.group is not dist.group.WORLD else r

    def exchange_planes(self, buf, n_own, glo, ghi, width):
        """Fill the ghost planes of `buf` (planes: [glo ghosts | n_own owned | ghi ghosts]) with the
        neighbours' outermost `width` owned planes."""
        if self.size == 1:
            return
        ops = []
        lo_nb, hi_nb = self.rank - 1, self.rank + 1
        if lo_nb >= 0:
            assert glo == width and n_own >= width
            ops.append(dist.P2POp(dist.isend, buf[glo:glo + width], self._peer(lo_nb), self.group))
            ops.append(dist.P2POp(dist.irecv, buf[0:glo], self._peer(lo_nb), self.group))
        if hi_nb < self.size:
            assert ghi == width and n_own >= width
            ops.append(dist.P2POp(dist.isend, buf[glo + n_own - width:glo + n_own],
                                  self._peer(hi_nb), self.group))
            ops.append(dist.P2POp(dist.irecv, buf[glo + n_own:glo + n_own + ghi],
                                  self._peer(hi_nb), self.group))
        with profiling.region("halo_exchange", 2 * 8 * width * buf[0].numel()):
            for req in dist.batch_isend_irecv(ops):
                req.wait()

    def exchange(self, v):
        """Halo exchange of a StencilVector (p planes per neighbour): peer stores over NVLink when
        the vector lives in the peer arena on every rank, NCCL send/recv otherwise."""
        V = v.space
        if self.size > 1 and self._p2p_ready(v):
            self._exchange_p2p(v)
            return
        self.exchange_planes(v._buf, V.local_shape[0], V.glo, V.ghi, V.pads[0])

    def _p2p_ready(self, v):
        """Decided ONCE per vector, collectively (first exchange): every rank holds it at the same
        arena position."""
        st = v.__dict__.get("_p2p")
        if st is None:
            key = v.__dict__.get("_arena_key") if self.arena is not None else None
            if key is None:
                # arena blocks are handed out in the same order and size on every rank (SPMD), so a
                # vector outside the arena here is outside it everywhere: NCCL, no handshake
                v.__dict__["_p2p"] = False
                return False
            keys = [None] * self.size
            dist.all_gather_object(keys, key, group=self.group)
            st = v.__dict__["_p2p"] = bool(key is not None and all(k == key for k in keys))
        return st

    def _exchange_p2p(self, v):
        from . import _lib
        V = v.space
        a = self.arena
        w = V.pads[0]
        per = int(np.prod(V.pitched_shape[1:]))              # doubles per plane
        i, off = v._arena_key
        ch, fl = a.chunks[i], a.chunks[0]
        base = v._buf.data_ptr()
        assert base == ch["base"] + off
        o = p2p_offsets(V.npts[0], self.size, self.rank, w)
        assert o["n_own"] == V.local_shape[0] and o["glo"] == V.glo and o["ghi"] == V.ghi
        src_lo = dst_lo = src_hi = dst_hi = lo_f = hi_f = None
        if o["src_lo"] is not None:
            src_lo = base + 8 * per * o["src_lo"]
            dst_lo = ch["lo"] + off + 8 * per * o["dst_lo"]              # its upper ghost planes
            lo_f = fl["lo"]
        if o["src_hi"] is not None:
            src_hi = base + 8 * per * o["src_hi"]
            dst_hi = ch["hi"] + off + 8 * per * o["dst_hi"]              # its lower ghost planes
            hi_f = fl["hi"]
        with profiling.region("halo_exchange", 2 * 8 * w * per):
            _lib.check(a.L.poms_halo_exchange_p2p(src_lo, dst_lo, src_hi, dst_hi, w * per, fl["base"],
                                                  lo_f, hi_f, torch.cuda.current_stream().cuda_stream),
                       "poms_halo_exchange_p2p")

    def exchange_async(self, v):
        """The same exchange on the communication stream, ordered after the work already queued on
        the current stream; returns the event the consumer of the ghost planes must wait for."""
        if self._comm_stream is None:
            self._comm_stream = torch.cuda.Stream(device=self.device)
        cs = self._comm_stream
        cs.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(cs):
            self.exchange(v)
            ev = cs.record_event()
        return ev

    def gather_planes(self, own, table, need):
        """Planes need[0]..need[1] (global numbering, inclusive) of a plane-partitioned array of
        which this rank holds `own` = planes table[rank]; the missing ones come from the lower /
        upper neighbour.  `need_of(r)` must be the same function on every rank: pass the table of
        needs as need = [(lo, hi) for every rank]."""
        s, e = table[self.rank]
        lo, hi = need[self.rank]
        assert lo <= s and hi >= e
        out = own.new_empty((hi - lo + 1,) + tuple(own.shape[1:]))
        out[s - lo:e - lo + 1] = own
        if self.size == 1:
            assert lo == s and hi == e
            return out
        ops = []
        r = self.rank
        if r > 0:
            ps, pe = table[r - 1]
            plo, phi = need[r - 1]
            assert lo >= ps, "needs planes beyond the direct neighbour"
            if lo < s:                       # I need [lo, s-1] from below
                ops.append(dist.P2POp(dist.irecv, out[0:s - lo], self._peer(r - 1), self.group))
            if phi > pe:                     # the lower neighbour needs [pe+1, phi] from me
                assert phi <= e
                ops.append(dist.P2POp(dist.isend, own[0:phi - pe].contiguous(),
                                      self._peer(r - 1), self.group))
        if r < self.size - 1:
            ns, ne = table[r + 1]
            nlo, nhi = need[r + 1]
            assert hi <= ne, "needs planes beyond the direct neighbour"
            if hi > e:
                ops.append(dist.P2POp(dist.irecv, out[e - lo + 1:], self._peer(r + 1), self.group))
            if nlo < ns:                     # the upper neighbour needs [nlo, ns-1] from me
                assert nlo >= s
                ops.append(dist.P2POp(dist.isend, own[nlo - s:].contiguous(),
                                      self._peer(r + 1), self.group))
        if ops:
            with profiling.region("gather_planes", 0):
                for req in dist.batch_isend_irecv(ops):
                    req.wait()
        return out

    def allgather_planes(self, own, table):
        """Full array (all planes, every rank) from its plane-partitioned pieces."""
        if self.size == 1:
            return own
        n = table[-1][1] + 1
        full = own.new_empty((n,) + tuple(own.shape[1:]))
        sizes = [e - s + 1 for s, e in table]
        if len(set(sizes)) == 1:
            dist.all_gather_into_tensor(full, own.contiguous(), group=self.group)
            return full
        # uneven blocks: pad to the largest one (the gathered levels are small)
        mx = max(sizes)
        padded = own.new_zeros((mx,) + tuple(own.shape[1:]))
        padded[:own.shape[0]] = own
        stack = own.new_empty((self.size * mx,) + tuple(own.shape[1:]))
        dist.all_gather_into_tensor(stack, padded, group=self.group)
        for q, (s, e) in enumerate(table):
            full[s:e + 1] = stack[q * mx:q * mx + (e - s + 1)]
        return full


def ghost_view_range(table, need, rank, width, glo, n_own):
    """Planes need[rank] = [lo, hi] of a slab-partitioned array as a slice [a, b) of the rank's
    storage [glo ghost planes | n_own owned | ghost planes] once a halo exchange of `width` planes
    has filled the ghosts -- or None when SOME rank needs more than `width` planes of a neighbour
    (the decision must be the same on every rank: the exchange is collective)."""
    if any(ts - tl > width or th - te > width for (ts, te), (tl, th) in zip(table, need)):
        return None
    s, e = table[rank]
    lo, hi = need[rank]
    assert s - lo <= glo
    return glo - (s - lo), glo + n_own + (hi - e)


# ==========================================================================================
# plan of the axis-1 grid transfer between two slab-partitioned (or partitioned -> replicated) levels
# ==========================================================================================
def _row_extent(start, coef):
    """Per row: first and last column holding a non-zero coefficient."""
    W = coef.shape[1]
    nz = coef != 0.0
    first = np.where(nz.any(axis=1), nz.argmax(axis=1), 0)
    last = np.where(nz.any(axis=1), W - 1 - nz[:, ::-1].argmax(axis=1), 0)
    return start + first, start + last


def slab_transfer_plan(st, cf, nc, size, coarse_distributed):
    """Host-side plan for applying P1 (rows (st, cf): fine row i <- coarse cols st[i]+w) and
    R1 = P1^T along the partitioned axis.  For every rank q:
      need_f[q]  fine planes [lo, hi] its owned coarse rows read (restriction),
      R0[q]      (start_local, coef) of those rows relative to plane lo,
      need_c[q]  coarse planes its owned fine rows read (prolongation; None when the coarse level
                 is replicated and every rank already holds all of them),
      P0[q]      (start_local, coef, n_in)."""
    from . import bsplines as bs
    nf = len(st)
    stt, cft = bs.rows_transpose(st, cf, nc)
    tf = [block_bounds(nf, size, q) for q in range(size)]
    tc = [block_bounds(nc, size, q) for q in range(size)]
    f_first, f_last = _row_extent(stt, cft)
    c_first, c_last = _row_extent(st, cf)
    plan = dict(tf=tf, tc=tc, need_f=[], need_c=[] if coarse_distributed else None, R0=[], P0=[])
    for q in range(size):
        cs, ce = tc[q]
        fs, fe = tf[q]
        lo = int(min(f_first[cs:ce + 1].min(), fs))
        hi = int(max(f_last[cs:ce + 1].max(), fe))
        plan["need_f"].append((lo, hi))
        plan["R0"].append((stt[cs:ce + 1] - lo, cft[cs:ce + 1], hi - lo + 1))
        if coarse_distributed:
            lo = int(min(c_first[fs:fe + 1].min(), cs))
            hi = int(max(c_last[fs:fe + 1].max(), ce))
            plan["need_c"].append((lo, hi))
            plan["P0"].append((st[fs:fe + 1] - lo, cf[fs:fe + 1], hi - lo + 1))
        else:
            plan["P0"].append((st[fs:fe + 1], cf[fs:fe + 1], nc))
    return plan


# ==========================================================================================
# SPIKE: banded solve along the partitioned axis
# ==========================================================================================
class SpikeSetup:
    """Host-side setup of the partitioned solve T x = y for a banded matrix T (half-bandwidth q)
    whose rows are split into `table` blocks.  With x~_r = T_rr^-1 y_r,
        x_r = x~_r - W_r x_{r-1}^{bot} - V_r x_{r+1}^{top}
    (top / bot = first / last q entries).  Stacking the top and bottom q rows of every block gives a
    2qG x 2qG system S z = z~ that is the same for every line; Sinv is applied as a dense
    contraction.  V_r, W_r decay exponentially away from the interface for the diagonally dominant
    SPD bands of this code (mass, GLT): rows below `tol` relative are dropped."""

    def __init__(self, band, table, tol=1e-17):
        from . import bsplines as bs
        band = np.asarray(band, dtype=np.float64)
        p = (band.shape[1] - 1) // 2
        q = p
        while q > 0 and not band[:, p - q].any() and not band[:, p + q].any():
            q -= 1
        self.q = q
        band = band[:, p - q:p + q + 1]
        n = band.shape[0]
        G = len(table)
        self.G, self.n, self.table = G, n, table
        self.local_bands, self.V, self.W = [], [], []
        S = np.eye(2 * q * G)
        for r, (s, e) in enumerate(table):
            nl = e - s + 1
            assert nl >= 2 * q, "a slab needs at least 2q planes"
            lb = band[s:e + 1].copy()
            # entries that couple to other blocks are not part of the diagonal block
            for k in range(-q, q + 1):
                i = np.arange(nl)
                out = (i + k < 0) | (i + k >= nl)
                lb[out, k + q] = 0.0
            self.local_bands.append(lb)
            from scipy.linalg import solve_banded
            ab = np.zeros((2 * q + 1, nl))
            for k in range(-q, q + 1):
                i = np.arange(max(0, -k), min(nl, nl - k))
                ab[q - k, i + k] = lb[i, k + q]
            B = np.zeros((nl, q))      # coupling to the first q entries of block r+1
            C = np.zeros((nl, q))      # coupling to the last q entries of block r-1
            for i in range(nl):
                for k in range(-q, q + 1):
                    j = s + i + k
                    if j > e and j < n:
                        B[i, j - (e + 1)] = band[s + i, k + q]
                    if j < s and j >= 0:
                        C[i, j - (s - q)] = band[s + i, k + q]
            V = solve_banded((q, q), ab, B) if r < G - 1 else np.zeros((nl, q))
            W = solve_banded((q, q), ab, C) if r > 0 else np.zeros((nl, q))
            self.V.append(V)
            self.W.append(W)
            # reduced system rows of block r: [top_r (q) ; bot_r (q)]
            o = 2 * q * r
            if r > 0:
                S[o:o + q, o - q:o] += W[:q]                 # top_r  + W^top  bot_{r-1}
                S[o + q:o + 2 * q, o - q:o] += W[nl - q:]    # bot_r  + W^bot  bot_{r-1}
            if r < G - 1:
                S[o:o + q, o + 2 * q:o + 3 * q] += V[:q]             # top_r + V^top top_{r+1}
                S[o + q:o + 2 * q, o + 2 * q:o + 3 * q] += V[nl - q:]
        self.Sinv = np.linalg.inv(S)
        # truncated spikes: number of planes from the interface where they are above rounding
        def depth(M, from_top):
            a = np.abs(M).max(axis=1)
            a = a if from_top else a[::-1]
            nz = np.nonzero(a > tol * max(a.max(), 1e-300))[0]
            return int(nz[-1]) + 1 if len(nz) else 0
        self.mW = [depth(W, True) for W in self.W]
        self.mV = [depth(V, False) for V in self.V]


def spike_solve_host(setup, y_blocks):
    """Reference implementation of the partitioned solve on the host (tests): y_blocks[r] is the
    (n_r, m) block of right-hand sides of slab r; returns the x blocks."""
    from scipy.linalg import solve_banded
    q, G = setup.q, setup.G
    xt = []
    for r in range(G):
        lb = setup.local_bands[r]
        nl = lb.shape[0]
        ab = np.zeros((2 * q + 1, nl))
        for k in range(-q, q + 1):
            i = np.arange(max(0, -k), min(nl, nl - k))
            ab[q - k, i + k] = lb[i, k + q]
        xt.append(solve_banded((q, q), ab, y_blocks[r]))
    zt = np.concatenate([np.concatenate([x[:q], x[-q:]]) for x in xt])
    z = setup.Sinv @ zt
    out = []
    for r in range(G):
        x = xt[r].copy()
        o = 2 * q * r
        if r > 0:
            x -= setup.W[r] @ z[o - q:o]
        if r < G - 1:
            x -= setup.V[r] @ z[o + 2 * q:o + 3 * q]
        out.append(x)
    return out


def _spike_for(lu, V):
    """SPIKE setup + device operators of one axis-1 factor for the partition of space V (cached)."""
    from .kron_product import BandLU
    from .mg import _AxisOp
    slab = V.slab
    # the cache lives ON the factor object: an id()-keyed module cache would hand a stale setup to a
    # new BandLU that happens to reuse the id of a freed one
    cache = lu.__dict__.setdefault("_spike", {})
    key = (V.npts[0], slab.size, slab.rank, str(V.device))
    ent = cache.get(key)
    if ent is not None:
        return ent
    if getattr(lu, "band", None) is None:
        raise NotImplementedError(
            "slab-partitioned Kronecker solve needs the band matrix itself (BandLU.from_band); "
            "a bare dgbtrf factorisation [A_bnd, la, ua, piv] cannot be re-partitioned")
    table = slab.table(V.npts[0])
    st = SpikeSetup(lu.band, table)
    q, G, r = st.q, st.G, slab.rank
    dev = V.device
    ent = {"setup": st, "q": q, "local_lu": BandLU.from_band(st.local_bands[r], dev)}
    nl_all = [e - s_ + 1 for s_, e in table]
    # Thick slabs: the spikes have decayed below rounding before they reach the far side of a slab
    # (W^bot = V^top = 0 to 1e-17), so the reduced system splits into one independent 2q x 2q
    # system per interface that only involves the two slabs touching it: a q-plane neighbour
    # exchange replaces the all-gather of every interface plane.
    local = all(st.mW[k] <= nl_all[k] - q and st.mV[k] <= nl_all[k] - q for k in range(G))
    ent["local"] = local
    if local:
        def iface_inverse(lo):   # interface between slabs lo and lo+1: unknowns [bot_lo ; top_lo+1]
            nlo = nl_all[lo]
            S2 = np.eye(2 * q)
            S2[:q, q:] = st.V[lo][nlo - q:]          # bot_lo + V^bot_lo top_{lo+1}
            S2[q:, :q] = st.W[lo + 1][:q]            # top_{lo+1} + W^top_{lo+1} bot_lo
            return np.linalg.inv(S2)
        if r > 0:       # need bot_{r-1}: first q rows of the inverse of interface (r-1, r)
            ent["red_lo"] = _AxisOp(np.zeros(q, dtype=np.int32), iface_inverse(r - 1)[:q], 2 * q, dev)
        if r < G - 1:   # need top_{r+1}: last q rows of the inverse of interface (r, r+1)
            ent["red_hi"] = _AxisOp(np.zeros(q, dtype=np.int32), iface_inverse(r)[q:], 2 * q, dev)
    # rows of Sinv this rank needs: bot_{r-1} and top_{r+1}
    rows = []
    if r > 0:
        rows += list(range(2 * q * (r - 1) + q, 2 * q * (r - 1) + 2 * q))
    if r < G - 1:
        rows += list(range(2 * q * (r + 1), 2 * q * (r + 1) + q))
    if rows:
        ent["reduce"] = _AxisOp(np.zeros(len(rows), dtype=np.int32), st.Sinv[rows], 2 * q * G, dev)
    nl = table[r][1] - table[r][0] + 1
    if r > 0 and st.mW[r] > 0:
        m = st.mW[r]
        ent["corrW"] = (_AxisOp(np.zeros(m, dtype=np.int32), -st.W[r][:m], q, dev), m)
    if r < G - 1 and st.mV[r] > 0:
        m = st.mV[r]
        ent["corrV"] = (_AxisOp(np.zeros(m, dtype=np.int32), -st.V[r][nl - m:], q, dev), m)
    cache[key] = ent
    return ent


def kron_solve_bnd_slab(factors, Y, X=None, only_axis1=False):
    """X = (A_1 (x) .. (x) A_d)^-1 Y on a slab-partitioned space: SPIKE along axis 1, plain local
    line solves along the other axes (kron_product.kron_solve_bnd is the one-GPU version)."""
    from .kron_product import _solve_axis, BandLU
    from .stencil import StencilVector
    from . import profiling
    V = Y.space
    slab = V.slab
    if X is None:
        X = StencilVector(V)
    lus = list(factors)
    for f in (lus[:1] if only_axis1 else lus):
        if not isinstance(f, BandLU):
            raise NotImplementedError("slab-partitioned solve takes BandLU factors")
    sp = _spike_for(lus[0], V)
    q = sp["q"]
    shape = tuple(V.local_shape)
    n1 = shape[0]
    ld = V.ld
    # 1. local diagonal-block solves along axis 1
    with profiling.region("band_solve_axis1", 16 * V.local_size):
        _solve_axis(sp["local_lu"], Y, X, 0)
    if slab.size > 1 and sp["local"]:
        with profiling.region("spike_interface", 0, launches=4):
            r, G = slab.rank, slab.size
            glo = q if r > 0 else 0
            ghi = q if r < G - 1 else 0
            tail = tuple(X.flat.shape[1:])
            # 2. [from below | my top | my bottom | from above]: one q-plane neighbour exchange
            buf = torch.empty((glo + 2 * q + ghi,) + tail, dtype=X.flat.dtype, device=X.flat.device)
            buf[glo:glo + q] = X.flat[:q]
            buf[glo + q:glo + 2 * q] = X.flat[n1 - q:]
            slab.exchange_planes(buf, 2 * q, glo, ghi, q)
            # 3./4. per-interface 2q x 2q solves and the truncated spike corrections
            if r > 0:
                zprev = torch.empty((q,) + tail, dtype=buf.dtype, device=buf.device)
                sp["red_lo"].apply(buf[0:2 * q], zprev, (2 * q,) + shape[1:], ld, ld, 0)
                if "corrW" in sp:
                    cop, m = sp["corrW"]
                    cop.apply(zprev, X.flat[:m], (q,) + shape[1:], ld, ld, 0, accumulate=True)
            if r < G - 1:
                znext = torch.empty((q,) + tail, dtype=buf.dtype, device=buf.device)
                sp["red_hi"].apply(buf[glo + q:glo + 3 * q], znext, (2 * q,) + shape[1:], ld, ld, 0)
                if "corrV" in sp:
                    cop, m = sp["corrV"]
                    cop.apply(znext, X.flat[n1 - m:], (q,) + shape[1:], ld, ld, 0, accumulate=True)
    elif slab.size > 1:
        with profiling.region("spike_interface", 0, launches=3):
            # 2. interface planes of every slab -> everyone
            zl = torch.cat([X.flat[:q], X.flat[n1 - q:]], dim=0).contiguous()
            zt = torch.empty((2 * q * slab.size,) + tuple(zl.shape[1:]), dtype=zl.dtype,
                             device=zl.device)
            dist.all_gather_into_tensor(zt, zl, group=slab.group)
            # 3. the rows of the reduced solution this slab needs
            if "reduce" in sp:
                op = sp["reduce"]
                z = torch.empty((op.n_out,) + tuple(zl.shape[1:]), dtype=zl.dtype, device=zl.device)
                op.apply(zt, z, (2 * q * slab.size,) + shape[1:], ld, ld, 0)
                k = 0
                # 4. truncated spike corrections near the two interfaces
                if slab.rank > 0:
                    zprev = z[k:k + q]
                    k += q
                    if "corrW" in sp:
                        cop, m = sp["corrW"]
                        cop.apply(zprev, X.flat[:m], (q,) + shape[1:], ld, ld, 0, accumulate=True)
                if slab.rank < slab.size - 1:
                    znext = z[k:k + q]
                    if "corrV" in sp:
                        cop, m = sp["corrV"]
                        cop.apply(znext, X.flat[n1 - m:], (q,) + shape[1:], ld, ld, 0,
                                  accumulate=True)
    if only_axis1:
        return X
    src = X
    for ax in range(1, len(lus)):
        with profiling.region("band_solve_axis%d" % (ax + 1), 16 * V.local_size):
            _solve_axis(lus[ax], src, X, ax)
    return X
