// emu_stencil3d.cpp -- TEST INFRASTRUCTURE ONLY (see cuda_emu.h): poms_stencil_matvec_3d (full 3-D stencil,
// spl's StencilMatrix.dot in 3-D) and poms_color_add (two-colour Jacobi half sweep) of poms_extra.cu,
// rewritten for g++ by make_emu_source.py into stencil3d_emu.cuh.  Exactly sized buffers.
//   emu_stencil3d <in> <out>
// in:  int32 header (16): {op, n1, n2, n3, p1, p2, p3, epilogue, has_dot, glo, ghi, ld, off, colour, has_b, 0},
//      fp64 omega, then
//      op 0: x ((glo + n1 + ghi) * n2 * ld), b (n1 * n2 * ld, when has_b), S (n1 n2 n3 (2p1+1)(2p2+1)(2p3+1))
//      op 1: x (n1 * n2 * ld), d (same)
// out: int32 status, fp64 dot, y (n1 * n2 * ld)   [op 1: the updated x]
#define POMS_HOST_EMU 1
#include "cuda_emu.h"
#include "poms_b200.h"
#define POMS_WS_HEADER 256           // as in poms_kernels.cu (outside the extracted section)
#define POMS_MAX_PARTIALS 65536
#include "stencil3d_emu.cuh"

#include <cstdlib>
#include <memory>

static double* read_n(FILE* f, size_t n) {
    double* a = new double[n];
    if (fread(a, 8, n, f) != n) exit(3);
    return a;
}

int main(int argc, char** argv) {
    if (argc != 3) return 2;
    FILE* f = fopen(argv[1], "rb");
    if (!f) return 2;
    int32_t h[16];
    double omega;
    if (fread(h, 4, 16, f) != 16 || fread(&omega, 8, 1, f) != 1) return 3;
    const int op = h[0], n1 = h[1], n2 = h[2], n3 = h[3], p1 = h[4], p2 = h[5], p3 = h[6], epi = h[7], has_dot = h[8];
    const int glo = h[9], ghi = h[10], ld = h[11], off = h[12], colour = h[13], has_b = h[14];
    const int64_t pld = (int64_t)n2 * ld;
    const size_t nown = (size_t)n1 * pld;
    FILE* o = fopen(argv[2], "wb");
    int32_t rc = -99;
    double dot = 0.0;
    if (op == 0) {
        double* x = read_n(f, (size_t)(glo + n1 + ghi) * pld);
        double* b = has_b ? read_n(f, nown) : nullptr;
        double* S = read_n(f, (size_t)n1 * n2 * n3 * (2 * p1 + 1) * (2 * p2 + 1) * (2 * p3 + 1));
        double* y = new double[nown];
        for (size_t i = 0; i < nown; ++i) y[i] = 0.0;
        const size_t wsn = POMS_WS_HEADER + (size_t)POMS_MAX_PARTIALS * 8;
        std::unique_ptr<unsigned char[]> ws(new unsigned char[wsn]);
        memset(ws.get(), 0, wsn);
        rc = poms_stencil_matvec_3d(x + (size_t)glo * pld, y, b, S, n1, n2, n3, ld, pld, glo, ghi, p1, p2, p3, epi, omega,
                                    has_dot ? &dot : nullptr, ws.get(), nullptr);
        fwrite(&rc, 4, 1, o);
        fwrite(&dot, 8, 1, o);
        fwrite(y, 8, nown, o);
        delete[] x; delete[] b; delete[] S; delete[] y;
    } else {
        double* x = read_n(f, nown);
        double* d = read_n(f, nown);
        rc = poms_color_add(x, d, n1, n2, n3, ld, pld, off, colour, nullptr);
        fwrite(&rc, 4, 1, o);
        fwrite(&dot, 8, 1, o);
        fwrite(x, 8, nown, o);
        delete[] x; delete[] d;
    }
    if (rc != 0) fprintf(stderr, "status %d: %s\n", rc, g_err);
    fclose(f);
    fclose(o);
    return 0;
}
