"""TEST INFRASTRUCTURE ONLY: cut a kernel section out of poms_b200/csrc/poms_kernels.cu and rewrite the
CUDA-only syntax so that g++ can compile it over cuda_emu.h:
    kernel<T...><<<grid, block, smem, stream>>>(args)  ->  EMU_LAUNCH_EX((kernel<T...>), grid, block, smem, stream, args)
    extern __shared__ double name[];                      ->  static double name[EMU_DYN_SMEM_DOUBLES];
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
    launch = re.compile(r"([A-Za-z_]\w*(?:<[^<>;()]*>)?)<<<([^;]*?)>>>\(([^;()]*)\)")
    src, n = launch.subn(lambda m: "EMU_LAUNCH_EX((%s), %s, %s)" % (m.group(1), m.group(2), m.group(3)), src)
    assert "<<<" not in src, "unconverted launch"
    src = re.sub(r"extern\s+__shared__\s+(\w+)\s+(\w+)\[\];", r"static \1 \2[EMU_DYN_SMEM_DOUBLES];", src)
    return src, n


def band_solve_section():
    src = section(os.path.join(CSRC, "poms_kernels.cu"), "// K4: dgbtrs along one axis",
                  "// K5: per-axis sparse row gather")
    return to_host(src)


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


if __name__ == "__main__":
    s, n = band_solve_section()
    print(s[:400])
    print("...", n, "launches converted,", len(s.splitlines()), "lines")
