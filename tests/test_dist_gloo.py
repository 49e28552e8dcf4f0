"""Multi-rank host logic on CPU: world_size-2 (and 3) gloo process groups exercise the slab
partition, the halo exchange, the neighbour plane gathers of the grid transfer, the all-gather of
the gathered level and the scalar all-reduce with CPU tensors; the SPIKE partitioned banded solve
is checked against a dense solve.  (The CUDA kernels themselves are covered by -m gpu tests.)"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from poms_b200 import bsplines as bs
from poms_b200.dist import (Slab, block_bounds, SpikeSetup, spike_solve_host, slab_transfer_plan,
                             ghost_view_range)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _rows_apply(start, coef, X):
    out = np.zeros((coef.shape[0],) + X.shape[1:])
    for w in range(coef.shape[1]):
        j = start + w
        ok = (j >= 0) & (j < X.shape[0]) & (coef[:, w] != 0.0)
        out[ok] += coef[ok, w][:, None] * X[j[ok]]
    return out


def _worker(rank, world, port, p, N, coarse_distributed):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        slab = Slab(dist.group.WORLD, torch.device("cpu"))
        assert slab.rank == rank and slab.size == world
        rng = np.random.default_rng(42)
        nf, nc, m = N + p, N // 2 + p, 5
        # ---- partition covers the range exactly once
        tab = slab.table(nf)
        assert tab[0][0] == 0 and tab[-1][1] == nf - 1
        assert all(tab[i][1] + 1 == tab[i + 1][0] for i in range(world - 1))
        # ---- halo exchange of p planes
        Xg = rng.standard_normal((nf, m))
        s, e = tab[rank]
        glo = p if rank > 0 else 0
        ghi = p if rank < world - 1 else 0
        buf = torch.zeros((glo + e - s + 1 + ghi, m), dtype=torch.float64)
        buf[glo:glo + e - s + 1] = torch.from_numpy(Xg[s:e + 1])
        slab.exchange_planes(buf, e - s + 1, glo, ghi, p)
        assert np.array_equal(buf.numpy(), Xg[s - glo:e + 1 + ghi])
        # ---- all-reduce of a scalar (StencilVector.dot)
        t = torch.tensor([float(np.dot(Xg[s:e + 1, 0], Xg[s:e + 1, 1]))], dtype=torch.float64)
        slab.allreduce_sum(t)
        assert abs(t.item() - np.dot(Xg[:, 0], Xg[:, 1])) < 1e-12
        # ---- grid transfer along the partitioned axis
        Tf, Tc = bs.make_open_knots(p, nf), bs.make_open_knots(p, nc)
        st, cf, _ = bs.knot_insertion_rows(Tc, Tf, p)
        P1 = bs.rows_to_dense(st, cf, nc)
        plan = slab_transfer_plan(st, cf, nc, world, coarse_distributed)
        tf, tc = plan["tf"], plan["tc"]
        assert tf == tab
        own_f = torch.from_numpy(np.ascontiguousarray(Xg[s:e + 1]))
        planes = slab.gather_planes(own_f, tf, plan["need_f"]).numpy()
        lo, hi = plan["need_f"][rank]
        assert np.array_equal(planes, Xg[lo:hi + 1])
        # the same planes as a VIEW of the halo-exchanged storage (what DistTransfer uses when the
        # neighbours' planes lie within the ghost width; no copy of the slab)
        for width in (p, p + 1, 1):
            g_lo = width if rank > 0 else 0
            g_hi = width if rank < world - 1 else 0
            rngv = ghost_view_range(tf, plan["need_f"], rank, width, g_lo, e - s + 1)
            fits = all(ts - tl <= width and th - te <= width
                       for (ts, te), (tl, th) in zip(tf, plan["need_f"]))
            assert (rngv is not None) == fits
            if rngv is not None:
                gb = torch.zeros((g_lo + e - s + 1 + g_hi, m), dtype=torch.float64)
                gb[g_lo:g_lo + e - s + 1] = own_f
                slab.exchange_planes(gb, e - s + 1, g_lo, g_hi, width)
                assert np.array_equal(gb[rngv[0]:rngv[1]].numpy(), planes)
        r0s, r0c, r0n = plan["R0"][rank]
        assert r0n == hi - lo + 1
        rc_own = _rows_apply(r0s, r0c, planes)
        rc_ref = P1.T @ Xg
        cs, ce = tc[rank]
        assert np.abs(rc_own - rc_ref[cs:ce + 1]).max() < 1e-13
        # gathered (replicated) coarse level
        full = slab.allgather_planes(torch.from_numpy(rc_own), tc).numpy()
        assert np.abs(full - rc_ref).max() < 1e-13
        # prolongation
        Eg = rng.standard_normal((nc, m))
        p0s, p0c, p0n = plan["P0"][rank]
        if coarse_distributed:
            own_c = torch.from_numpy(np.ascontiguousarray(Eg[cs:ce + 1]))
            cplanes = slab.gather_planes(own_c, tc, plan["need_c"]).numpy()
            clo, chi = plan["need_c"][rank]
            assert np.array_equal(cplanes, Eg[clo:chi + 1]) and p0n == chi - clo + 1
        else:
            cplanes = Eg
        xf_own = _rows_apply(p0s, p0c, cplanes)
        assert np.abs(xf_own - (P1 @ Eg)[s:e + 1]).max() < 1e-13
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,p,N,cdist", [(2, 3, 32, True), (2, 3, 32, False), (3, 2, 48, True),
                                              (2, 5, 64, True), (2, 1, 16, False)])
def test_slab_plumbing_gloo(world, p, N, cdist):
    mp.spawn(_worker, args=(world, _free_port(), p, N, cdist), nprocs=world, join=True)


@pytest.mark.parametrize("p,q,n,G", [(3, 5, 131, 4), (3, 3, 67, 2), (2, 3, 45, 3), (5, 9, 260, 8),
                                      (1, 1, 40, 2)])
def test_spike_partitioned_solve(p, q, n, G):
    band = bs.glt_band(p, n, degree=q) if q != p else bs.assemble_1d_bands(
        p, bs.make_open_knots(p, n))[0]
    table = [block_bounds(n, G, r) for r in range(G)]
    st = SpikeSetup(band, table)
    y = np.random.default_rng(0).standard_normal((n, 4))
    x = np.concatenate(spike_solve_host(st, [y[s:e + 1] for s, e in table]))
    xr = np.linalg.solve(bs.band_to_dense(band), y)
    assert np.abs(x - xr).max() < 1e-13 * np.abs(xr).max()
    # truncated spikes: the correction only touches planes near the interfaces
    assert max(st.mW) <= 60 and max(st.mV) <= 60


def test_peer_store_halo_offsets_reproduce_a_ghost_update():
    """Host logic of the peer-store halo exchange (dist.p2p_offsets): emulate the pushes of every rank
    on NumPy slabs of an uneven partition and compare the ghost planes with the global array."""
    import numpy as np
    from poms_b200.dist import p2p_offsets, block_bounds
    for n_glob, size, w in ((37, 4, 3), (515, 8, 4), (16, 2, 1), (131, 3, 2)):
        G = np.arange(n_glob * 5, dtype=float).reshape(n_glob, 5)
        offs = [p2p_offsets(n_glob, size, r, w) for r in range(size)]
        slabs = []
        for r, o in enumerate(offs):
            s, e = block_bounds(n_glob, size, r)
            buf = np.full((o["glo"] + o["n_own"] + o["ghi"], 5), np.nan)
            buf[o["glo"]:o["glo"] + o["n_own"]] = G[s:e + 1]
            slabs.append(buf)
        for r, o in enumerate(offs):                      # every rank pushes to its neighbours
            if o["src_lo"] is not None:
                slabs[r - 1][o["dst_lo"]:o["dst_lo"] + w] = slabs[r][o["src_lo"]:o["src_lo"] + w]
            if o["src_hi"] is not None:
                slabs[r + 1][o["dst_hi"]:o["dst_hi"] + w] = slabs[r][o["src_hi"]:o["src_hi"] + w]
        for r, o in enumerate(offs):
            s, e = block_bounds(n_glob, size, r)
            want = G[s - o["glo"]:e + 1 + o["ghi"]]
            assert np.array_equal(slabs[r], want), (n_glob, size, w, r)
