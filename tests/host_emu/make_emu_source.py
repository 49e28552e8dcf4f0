"""TEST INFRASTRUCTURE ONLY: cut a kernel section out of poms_b200/csrc/poms_kernels.cu and rewrite the
CUDA-only syntax so that g++ can compile it over cuda_emu.h:
    kernel<T...><<<grid, block, smem, stream>>>(args)  ->  EMU_LAUNCH_EX((kernel<T...>), grid, block, smem, stream, args)
    extern __shared__ T name[];                           ->  T* const name = reinterpret_cast<T*>(emu_dyn_smem);
The product source is not modified; the output goes to a temporary directory of the test."""
import os
import re

CSRC = os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))),
                    "poms_b200", "csrc")


def section(path, start_marker, end_marker):
    text = open(path).read()
    a = text.index(start_marker)
    b = text.index(end_marker, a)
    a = text.rfind("\n", 0, text.rfind("\n", 0, a)) + 1        # include the rule line above the title
    b = text.rfind("\n", 0, text.rfind("\n", 0, b)) + 1
    return text[a:b]


def to_host(src):
    src = src.replace("\\\n", " ")                             # join macro continuation lines
    launch = re.compile(r"([A-Za-z_]\w*(?:<[^<>;()]*>)?)<<<([^;]*?)>>>\(((?:[^;()]|\([^;()]*\))*)\)")
    src, n = launch.subn(lambda m: "EMU_LAUNCH_EX((%s), %s, %s)" % (m.group(1), m.group(2), m.group(3)), src)
    assert "<<<" not in src, "unconverted launch"
    src = re.sub(r"extern\s+__shared__\s+(?:__align__\(\d+\)\s+)?((?:unsigned\s+)?\w+)\s+(\w+)\[\];",
                 r"\1* const \2 = reinterpret_cast<\1*>(emu_dyn_smem);", src)
    assert "extern __shared__" not in src
    return src, n


def band_solve_section():
    src = section(os.path.join(CSRC, "poms_kernels.cu"), "// K4: dgbtrs along one axis",
                  "// K5: per-axis sparse row gather")
    return to_host(src)


def axis_gather_section():
    """Per-axis sparse row gather (2-D transfers, fallback of the 3-D ones, slab-plan rows)."""
    text = open(os.path.join(CSRC, "poms_kernels.cu")).read()
    a = text.index("// K5: per-axis sparse row gather")
    a = text.rfind("\n", 0, text.rfind("\n", 0, a)) + 1
    b = text.index("// dense mat-vec for the replicated coarse solve")
    b = text.rfind("\n", 0, text.rfind("\n", 0, b)) + 1
    return to_host(text[a:b])


def dmma_section():
    """Dense per-axis contraction on the fp64 tensor cores (poms_extra.cu): the coarse solve of the bench.
    The one-instruction wrapper dmma8x8x4 (mma.sync.m8n8k4.f64) is removed; cuda_emu.h provides it with
    the fragment layout of the PTX ISA."""
    text = open(os.path.join(CSRC, "poms_extra.cu")).read()
    a = text.index("// Dense per-axis contraction on the fp64 tensor cores")
    a = text.rfind("\n", 0, text.rfind("\n", 0, a)) + 1
    b = text.index("// Full (non-separable) 3-D stencil mat-vec")
    b = text.rfind("\n", 0, text.rfind("\n", 0, b)) + 1
    return to_host(_strip_functions(text[a:b], ["dmma8x8x4"]))


def generic_mv3_section():
    """The common device helpers (deterministic grid reduction, shift-form partial sums, MV3) and the
    generic 3-D Kronecker mat-vec of translation unit 6 (the fallback for tiny or misaligned grids)."""
    text = open(os.path.join(CSRC, "poms_kernels.cu")).read()
    a = text.index("// deterministic grid reduction")
    a = text.rfind("\n", 0, text.rfind("\n", 0, a)) + 1
    b = text.index("#if POMS_TU == 6")
    c = text.index("\n", b) + 1
    d = text.index("#endif  // POMS_TU == 6")
    return to_host(text[a:b] + text[c:d])


def tu0_middle_section():
    """Translation unit 0 between the 2-D mat-vec and the band solves: round-1 2-D Kronecker mat-vec
    (tiny / misaligned grids) with its entry point, full 2-D stencil mat-vec, the CG vector algebra
    (cg_update, p_update, dot, axpby, ...), jacobi_first -- preceded by the common device helpers and
    the axis-1 chunk rule they use.  The TMA fast path of the 2-D entry is stubbed out by the harness
    (try_matvec2d_tma returns "not applicable")."""
    text = open(os.path.join(CSRC, "poms_kernels.cu")).read()
    a = text.index("// deterministic grid reduction")
    a = text.rfind("\n", 0, text.rfind("\n", 0, a)) + 1
    b = text.index("#if POMS_TU == 6")
    c = text.index("static int g_chunk_override = 0;")
    d = text.index("#endif", c)
    e = text.index("// K1: Kronecker banded mat-vec, 2-D.")
    e = text.rfind("\n", 0, text.rfind("\n", 0, e)) + 1
    f = text.index("// K4: dgbtrs along one axis")
    f = text.rfind("\n", 0, text.rfind("\n", 0, f)) + 1
    stub = ("static int try_matvec2d_tma(const MV2&, int, int, int, const double*, const int*, cudaStream_t) "
            "{ return 1; }\n")
    return to_host(text[a:b] + text[c:d] + stub + text[e:f])


def _strip_functions(src, names):
    """Remove the definitions of small `__device__ __forceinline__ void NAME(...) { ... }` wrappers
    (inline PTX): the emulation header provides their host versions."""
    for name in names:
        src, n = re.subn(r"__device__ __forceinline__ void %s\(.*?\n}\n" % name, "", src, flags=re.S)
        assert n == 1, (name, n)
    return src


def tma_mv3_section(degrees=(2, 3, 4)):
    """The hot path: TMA-staged 3-D Kronecker mat-vec (poms_matvec3d_tma.cuh + poms_matvec3d_v3.cuh)
    behind poms_kron_matvec_3d_dotv, with the generic kernel as its fallback.  The device section that
    build.py compiles once per degree (-DPOMS_TU=p) is instantiated textually per degree inside a
    namespace; the inline-PTX wrappers (mbarrier, cp.async.bulk.tensor) are removed, emu_tma.h provides
    them; degrees that are not built answer "bad argument"."""
    text = open(os.path.join(CSRC, "poms_kernels.cu")).read()
    tma = open(os.path.join(CSRC, "poms_matvec3d_tma.cuh")).read()
    v3 = open(os.path.join(CSRC, "poms_matvec3d_v3.cuh")).read().replace("#pragma once", "")
    # common helpers + generic kernel (fallback) + chunk rule
    a = text.index("// deterministic grid reduction")
    a = text.rfind("\n", 0, text.rfind("\n", 0, a)) + 1
    b = text.index("#if POMS_TU == 6")
    c = text.index("\n", b) + 1
    d = text.index("#endif  // POMS_TU == 6")
    ck = text.index("static int g_chunk_override = 0;")
    ckd = text.index("#endif", ck)
    out = text[a:b] + text[c:d] + text[ck:ckd]
    # poms_matvec3d_tma.cuh: declarations before the per-degree section
    h0 = tma.index("struct MV3T {")
    h1 = tma.index("#if POMS_TU >= 1 && POMS_TU <= 5")
    out += _strip_functions(tma[h0:h1], ["mbar_init", "mbar_expect_tx", "mbar_wait", "tma_load_3d"])
    d0 = tma.index("\n", h1) + 1
    d1 = tma.index("#endif  // POMS_TU in 1..5")
    dev = tma[d0:d1].replace('#include "poms_matvec3d_v3.cuh"', _strip_functions(v3, ["mbar_arrive"]))
    dev = re.sub(r'[ \t]*asm volatile\("fence[^\n]*\n', "", dev)
    assert "asm volatile" not in dev
    sig = ("(const CUtensorMap* tm3, const MV3T& g, int form, int epi, int variant, int ntiles, dim3 grid, "
           "cudaStream_t st)")
    for P in range(1, 6):
        if P in degrees:
            out += "namespace emu_tu%d {\n#undef POMS_MV3_BLOCKS\n%s\n}\n" % (P, re.sub(r"\bPOMS_TU\b", str(P), dev))
            out += ("int poms_mv3_tma_launch_p%d%s { return emu_tu%d::poms_mv3_tma_launch_p%d(tm3, g, form, epi, "
                    "variant, ntiles, grid, st); }\n" % (P, sig, P, P))
        else:
            out += 'int poms_mv3_tma_launch_p%d%s { return bad_arg(11, "degree not built in the emulation"); }\n' % (P, sig)
    # host side of the TMA path (tensor-map cache, dispatch) and the C entry points
    t0 = tma.index("#if POMS_TU == 0")
    t0 = tma.index("\n", t0) + 1
    t1 = tma.index("#endif  // POMS_TU == 0")
    out += tma[t0:t1]
    e0 = text.index('extern "C" int poms_kron_matvec_3d_ex(')
    e1 = text.index("// K1: Kronecker banded mat-vec, 2-D.")
    e1 = text.rfind("\n", 0, text.rfind("\n", 0, e1)) + 1
    out += text[e0:e1]
    return to_host(out)


def tma_mv2_section():
    """The 2-D fast path: warp-autonomous TMA mat-vec (poms_matvec2d_tma.cuh, translation unit 7) behind
    poms_kron_matvec_2d_ex, with the round-1 kernel as its fallback (plus everything else of the
    tu0_middle_section, which shares the entry point's neighbourhood)."""
    text = open(os.path.join(CSRC, "poms_kernels.cu")).read()
    tma3 = open(os.path.join(CSRC, "poms_matvec3d_tma.cuh")).read()
    tma2 = open(os.path.join(CSRC, "poms_matvec2d_tma.cuh")).read()
    a = text.index("// deterministic grid reduction")
    a = text.rfind("\n", 0, text.rfind("\n", 0, a)) + 1
    b = text.index("#if POMS_TU == 6")
    c = text.index("static int g_chunk_override = 0;")
    d = text.index("#endif", c)
    c2 = text.index("// axis-1 chunk of the 2-D TMA kernel")
    d2 = text.index("#endif", c2)
    out = text[a:b] + text[c:d] + text[c2:d2]
    # mbarrier wrappers (stripped: emu_tma.h) and the driver-entry-point / cache-key helpers of the 3-D header
    h0 = tma3.index("struct MV3T {")
    h1 = tma3.index("#if POMS_TU >= 1 && POMS_TU <= 5")
    out += _strip_functions(tma3[h0:h1], ["mbar_init", "mbar_expect_tx", "mbar_wait", "tma_load_3d"])
    k0 = tma3.index("#include <unordered_map>")
    k1 = tma3.index("static int get_tmap(")
    out += tma3[k0:k1]
    # 2-D header: declarations, device section (translation unit 7), host section
    p0 = tma2.index("struct MV2T {")
    p1 = tma2.index("#if POMS_TU == 7")
    out += tma2[p0:p1]
    d0 = tma2.index("\n", p1) + 1
    d1 = tma2.index("#endif  // POMS_TU == 7")
    dev = _strip_functions(tma2[d0:d1], ["tma_load_2d"])
    dev = re.sub(r'[ \t]*asm volatile\("fence[^\n]*\n', "", dev)
    assert "asm volatile" not in dev
    out += dev
    t0 = tma2.index("#if POMS_TU == 0")
    t0 = tma2.index("\n", t0) + 1
    t1 = tma2.index("#endif", t0)
    out += tma2[t0:t1]
    e = text.index("// K1: Kronecker banded mat-vec, 2-D.")
    e = text.rfind("\n", 0, text.rfind("\n", 0, e)) + 1
    f = text.index("// K4: dgbtrs along one axis")
    f = text.rfind("\n", 0, text.rfind("\n", 0, f)) + 1
    out += text[e:f]
    return to_host(out)


if __name__ == "__main__":
    s, n = band_solve_section()
    print(s[:400])
    print("...", n, "launches converted,", len(s.splitlines()), "lines")


def setup_section():
    """Device-side 1-D set-up kernels (poms_setup.cu): assembly by quadrature, knot-insertion rows, banded LU
    without pivoting -- everything below the include of the common header."""
    text = open(os.path.join(CSRC, "poms_setup.cu")).read()
    a = text.index("#define POMS_MAXP")
    return to_host(text[a:])


def stencil3d_section():
    """Full (non-separable) 3-D stencil mat-vec and the two-colour update (poms_extra.cu), preceded by the
    common device helpers (warp / block sums, deterministic grid reduction) of poms_kernels.cu."""
    text = open(os.path.join(CSRC, "poms_kernels.cu")).read()
    a = text.index("// deterministic grid reduction")
    a = text.rfind("\n", 0, text.rfind("\n", 0, a)) + 1
    b = text.index("#if POMS_TU == 6")
    extra = open(os.path.join(CSRC, "poms_extra.cu")).read()
    c = extra.index("// Full (non-separable) 3-D stencil mat-vec")
    c = extra.rfind("\n", 0, extra.rfind("\n", 0, c)) + 1
    return to_host(text[a:b] + extra[c:])


def halo_section():
    """Peer-store halo exchange (poms_extra.cu).  The two inline-PTX wrappers (st.release.sys / ld.acquire.sys)
    are removed; emu_halo.cpp provides them as sequentially consistent atomics on memory shared between processes."""
    text = open(os.path.join(CSRC, "poms_extra.cu")).read()
    a = text.index("// Halo exchange by peer stores.")
    a = text.rfind("\n", 0, text.rfind("\n", 0, a)) + 1
    b = text.index("// Dense per-axis contraction on the fp64 tensor cores")
    b = text.rfind("\n", 0, text.rfind("\n", 0, b)) + 1
    src = text[a:b]
    for name in ("st_release_sys", "ld_acquire_sys"):
        src, n = re.subn(r"__device__ __forceinline__ \w+ %s\(.*?\n}\n" % name, "", src, flags=re.S)
        assert n == 1, name
    assert "asm volatile" not in src
    return to_host(src)
