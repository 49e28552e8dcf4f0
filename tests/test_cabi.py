"""CPU-side checks of the C-ABI boundary: the shared library loads without a GPU, exports
every symbol include/poms_b200.h declares, and validates arguments before touching CUDA."""
import ctypes
import os
import re

import pytest

from conftest import ROOT
from poms_b200 import _lib


def _declared():
    src = open(os.path.join(ROOT, "include", "poms_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(poms_[a-z0-9_]+)\s*\(", src)))


def test_library_loads_and_exports_every_declared_symbol():
    assert os.path.exists(_lib.LIB_PATH), "run `python -m poms_b200.build` first"
    L = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared()
    assert len(names) >= 18
    for n in names:
        assert hasattr(L, n), "header declares %s but the library does not export it" % n
    # and the Python binding covers exactly the header
    assert sorted(_lib.EXPORTS) == names


def test_version_and_workspace():
    L = _lib.lib()
    assert L.poms_version() >= 100
    assert L.poms_workspace_bytes() >= 8 * 1024


def test_argument_validation_happens_before_cuda():
    L = _lib.lib()
    # null x -> -(argument index), no CUDA call is made
    rc = L.poms_kron_matvec_3d(None, None, None, 4, 4, 4, 4, 16, 0, 0, 3, 1, None, None, None, None,
                               None, None, 0, 0.0, None, None, None)
    assert rc == -1
    assert b"bad argument" in L.poms_last_error()
    rc = L.poms_kron_matvec_2d(1, 1, None, 4, 4, 2, 0, 0, 3, 1, 1, 1, 1, 1, 0, 0.0, None, None, None)
    assert rc == -6  # ld < n2
    rc = L.poms_band_solve_axis(1, 1, 1, 1, 8, 9, 1, 1, 8, 1, 1, None)
    assert rc == -6  # kl out of range
    with pytest.raises(_lib.PomsError):
        _lib.check(rc, "poms_band_solve_axis")


def test_fused_transfer_checks_rows_against_its_tiles_on_the_host():
    """poms_restrict_3d / poms_prolong_3d refuse rows that do not fit their static tiles (the caller
    then uses poms_axis_gather) -- decided from the host copies of the starts, before any launch."""
    import numpy as np
    from poms_b200 import bsplines as bs
    L = _lib.lib()
    p, Nc = 3, 16
    st, cf, _ = bs.knot_insertion_rows(bs.make_open_knots(p, Nc + p), bs.make_open_knots(p, 2 * Nc + p), p)
    stt, cft = bs.rows_transpose(st, cf, Nc + p)
    s_ok = np.ascontiguousarray(stt, dtype=np.int32)
    nf, nc = 2 * Nc + p, Nc + p
    args = lambda s1, s2, s3, W: (1, 1, nf, nf, nf, nf + 1, nf * (nf + 1), nc, nc, nc, nc + 1, nc * (nc + 1),
                                  1, 1, W, 1, 1, W, 1, 1, W, s1.ctypes.data, s2.ctypes.data,
                                  s3.ctypes.data, None)
    # rows as wide as the kernel allows at most 8 taps
    assert L.poms_restrict_3d(*args(s_ok, s_ok, s_ok, 9)) == -15
    # starts that jump by 8 per coarse row: a tile of 8 rows would need 61 fine rows (> 22)
    s_bad = np.ascontiguousarray(np.arange(nc) * 8, dtype=np.int32)
    assert L.poms_restrict_3d(*args(s_ok, s_bad, s_ok, cft.shape[1])) == -22
    assert b"poms_axis_gather" in L.poms_last_error()
    # decreasing starts along axis 1
    assert L.poms_restrict_3d(*args(np.ascontiguousarray(s_ok[::-1]), s_ok, s_ok, cft.shape[1])) == -22
    assert L.poms_restrict_3d(None, *args(s_ok, s_ok, s_ok, 5)[1:]) == -1


def test_v2_transfer_entries_check_rows_on_the_host():
    """poms_restrict_3d_v2 / poms_prolong_3d_v2 (big-level one-pass kernels): one row width on all
    axes, rows inside the tiles, monotone starts -- refused with a negative status before any launch,
    so that mg.Transfer falls back to the round-1 kernels / the per-axis gathers."""
    import numpy as np
    from poms_b200 import bsplines as bs
    L = _lib.lib()
    p, Nc = 3, 16
    st, cf, _ = bs.knot_insertion_rows(bs.make_open_knots(p, Nc + p), bs.make_open_knots(p, 2 * Nc + p), p)
    stt, cft = bs.rows_transpose(st, cf, Nc + p)
    sr = np.ascontiguousarray(stt, dtype=np.int32)
    sp = np.ascontiguousarray(st, dtype=np.int32)
    nf, nc = 2 * Nc + p, Nc + p
    dims = (nf, nf, nf, nf + 1, nf * (nf + 1), nc, nc, nc, nc + 1, nc * (nc + 1))
    WR, WP = cft.shape[1], cf.shape[1]
    rargs = lambda s1, s2, s3, W1, W2, W3: (1, 1) + dims + (1, 1, W1, 1, 1, W2, 1, 1, W3, s1.ctypes.data,
                                                          s2.ctypes.data, s3.ctypes.data, None)
    pargs = lambda s2, s3, W1, W2, W3: (1, 1) + dims + (1, 1, W1, 1, 1, W2, 1, 1, W3, s2.ctypes.data,
                                                      s3.ctypes.data, 1, None)
    assert L.poms_restrict_3d_v2(*rargs(sr, sr, sr, WR, WR - 1, WR)) == -15      # mixed widths
    assert L.poms_restrict_3d_v2(*rargs(sr, sr, sr, 8, 8, 8)) == -15            # wider than 7
    assert L.poms_prolong_3d_v2(*pargs(sp, sp, WP, WP, WP + 1)) == -15
    assert b"poms_prolong_3d" in L.poms_last_error()
    s_bad = np.ascontiguousarray(np.arange(nc) * 8, dtype=np.int32)
    assert L.poms_restrict_3d_v2(*rargs(sr, s_bad, sr, WR, WR, WR)) == -22      # tile of 8 rows > 22 fine rows
    assert L.poms_restrict_3d_v2(*rargs(np.ascontiguousarray(sr[::-1]), sr, sr, WR, WR, WR)) == -22
    s_wide = np.ascontiguousarray(np.arange(nf) * 2, dtype=np.int32)
    assert L.poms_prolong_3d_v2(*pargs(s_wide, sp, WP, WP, WP)) == -22          # 16 fine rows > 16 coarse rows
    assert L.poms_prolong_3d_v2(None, *pargs(sp, sp, WP, WP, WP)[1:]) == -1
    assert L.poms_restrict_3d_v2(1, None, *rargs(sr, sr, sr, WR, WR, WR)[2:]) == -1


def test_no_cpu_path():
    import torch
    from poms_b200.stencil import DeviceContext
    with pytest.raises(_lib.PomsError):
        DeviceContext.get(torch.device("cpu"))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "poms_b200")
    for f in os.listdir(pkg):
        if f.endswith(".py"):
            src = open(os.path.join(pkg, f)).read()
            assert "import oracle" not in src and "from oracle" not in src, f
            assert "poms_oracle" not in src, f
