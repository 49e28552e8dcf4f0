// emu_tu0.cpp -- TEST INFRASTRUCTURE ONLY (see cuda_emu.h): kernels of translation unit 0 of
// poms_kernels.cu between the 2-D mat-vec and the band solves (rewritten for g++ by make_emu_source.py
// into tu0_middle_emu.cuh), run on the host through their C entry points:
//   op 0  poms_kron_matvec_2d   (round-1 2-D kernel: tiny / misaligned grids; TMA fast path stubbed out)
//   op 1  poms_cg_update        op 2  poms_p_update      op 3  poms_dot      op 4  poms_axpby
//   op 5  poms_jacobi_first_2d  op 6  poms_cheb_update
//   op 8  poms_axpy_dev   op 9  poms_diag_scale   op 10  poms_jacobi_first_3d (n3 = h[10]; bands m1 k1 m2 k2 m3 k3, b)
//   op 7  poms_stencil_matvec_2d  (full 2-D stencil; header: p = p1, form = p2, h[10] / h[11] = ghost rows below / above;
//         arrays: S (n1 n2 (2p1+1)(2p2+1)), x ((glo + n1 + ghi) * ld), b (n1 * ld, if has_b))
//   emu_tu0 <in> <out>
// in:  int32 header (16): {op, n1, n2, ld, p, form, epi, has_b, has_dot, n (flat length), 0...}
//      fp64 scalars (4): {omega | a, num | b, den, c2}
//      op 0 / 5: bands m1 k1 m2 k2 (n_a * (2p+1)), x (n1 * ld), b (n1 * ld, if has_b)
//      op 1..4, 6: up to four flat arrays of length n
// out: int32 status, fp64 dot, then the result arrays (op 0/5: y; op 1: x, r; op 2: p; op 3: -; op 4: z;
//      op 6: x, d)
#define POMS_HOST_EMU 1
#include "cuda_emu.h"
#include "poms_b200.h"
#define POMS_WS_HEADER 256
#define POMS_MAX_PARTIALS 65536
#include "tu0_middle_emu.cuh"

#include <cstdlib>
#include <memory>

template <class T>
static std::unique_ptr<T[]> rd(FILE* f, size_t n) {
    std::unique_ptr<T[]> p(new T[n ? n : 1]);
    if (n && fread(p.get(), sizeof(T), n, f) != n) {
        fprintf(stderr, "short read\n");
        exit(3);
    }
    return p;
}

int main(int argc, char** argv) {
    if (argc != 3) return 2;
    FILE* f = fopen(argv[1], "rb");
    if (!f) return 2;
    auto h = rd<int32_t>(f, 16);
    auto sc = rd<double>(f, 4);
    const int op = h[0], n1 = h[1], n2 = h[2], ld = h[3], p = h[4], form = h[5], epi = h[6], has_b = h[7],
              has_dot = h[8];
    const size_t n = (size_t)h[9];
    const size_t wsn = POMS_WS_HEADER + (size_t)POMS_MAX_PARTIALS * 8;
    std::unique_ptr<unsigned char[]> ws(new unsigned char[wsn]);
    memset(ws.get(), 0, wsn);
    double dot = 0.0;
    int rc = -99;
    FILE* o = fopen(argv[2], "wb");
    auto put = [&](const double* a, size_t cnt) { fwrite(a, 8, cnt, o); };
    int32_t rc32 = 0;
    fwrite(&rc32, 4, 1, o);
    fwrite(&dot, 8, 1, o);
    if (op == 0 || op == 5) {
        const int W = 2 * p + 1;
        auto m1 = rd<double>(f, (size_t)n1 * W), k1 = rd<double>(f, (size_t)n1 * W);
        auto m2 = rd<double>(f, (size_t)n2 * W), k2 = rd<double>(f, (size_t)n2 * W);
        const size_t total = (size_t)n1 * ld;
        auto x = rd<double>(f, total);
        auto b = rd<double>(f, has_b ? total : 0);
        std::unique_ptr<double[]> y(new double[total]);
        for (size_t i = 0; i < total; ++i) y[i] = 0.0;
        if (op == 0)
            rc = poms_kron_matvec_2d(x.get(), y.get(), has_b ? b.get() : nullptr, n1, n2, ld, 0, 0, p, form,
                                     m1.get(), k1.get(), m2.get(), k2.get(), epi, sc[0], has_dot ? &dot : nullptr,
                                     ws.get(), nullptr);
        else
            rc = poms_jacobi_first_2d(y.get(), b.get(), n1, n2, ld, p, form, m1.get(), k1.get(), m2.get(), k2.get(),
                                      sc[0], has_dot ? &dot : nullptr, ws.get(), nullptr);
        put(y.get(), total);
    } else if (op == 10) {      // first Jacobi sweep in 3-D: extents n1, n2, h[10]; ld = h[3]
        const int n3 = h[10], W = 2 * p + 1;
        auto m1 = rd<double>(f, (size_t)n1 * W), k1 = rd<double>(f, (size_t)n1 * W);
        auto m2 = rd<double>(f, (size_t)n2 * W), k2 = rd<double>(f, (size_t)n2 * W);
        auto m3 = rd<double>(f, (size_t)n3 * W), k3 = rd<double>(f, (size_t)n3 * W);
        const size_t total = (size_t)n1 * n2 * ld;
        auto b = rd<double>(f, total);
        std::unique_ptr<double[]> y(new double[total]);
        for (size_t i = 0; i < total; ++i) y[i] = 0.0;
        rc = poms_jacobi_first_3d(y.get(), b.get(), n1, n2, n3, ld, (int64_t)n2 * ld, p, form, m1.get(), k1.get(), m2.get(),
                                  k2.get(), m3.get(), k3.get(), sc[0], has_dot ? &dot : nullptr, ws.get(), nullptr);
        put(y.get(), total);
    } else if (op == 7) {       // full 2-D stencil (spl StencilMatrix.dot): p1 = p, p2 = form, ghost rows h[10] / h[11]
        const int p2 = form, glo = h[10], ghi = h[11];
        auto S = rd<double>(f, (size_t)n1 * n2 * (2 * p + 1) * (2 * p2 + 1));
        auto x = rd<double>(f, (size_t)(glo + n1 + ghi) * ld);
        const size_t total = (size_t)n1 * ld;
        auto b = rd<double>(f, has_b ? total : 0);
        std::unique_ptr<double[]> y(new double[total]);
        for (size_t i = 0; i < total; ++i) y[i] = 0.0;
        rc = poms_stencil_matvec_2d(x.get() + (size_t)glo * ld, y.get(), has_b ? b.get() : nullptr, S.get(), n1, n2, ld,
                                    glo, ghi, p, p2, epi, sc[0], has_dot ? &dot : nullptr, ws.get(), nullptr);
        put(y.get(), total);
    } else {
        auto a0 = rd<double>(f, n), a1 = rd<double>(f, n), a2 = rd<double>(f, n), a3 = rd<double>(f, n);
        if (op == 1) {          // x, r, p, q ; alpha = num / den
            rc = poms_cg_update(a0.get(), a1.get(), a2.get(), a3.get(), (int64_t)n, &sc[1], &sc[2], &dot, ws.get(),
                                nullptr);
            put(a0.get(), n);
            put(a1.get(), n);
        } else if (op == 2) {   // p, s ; beta = num / den
            rc = poms_p_update(a0.get(), a1.get(), (int64_t)n, &sc[1], &sc[2], nullptr);
            put(a0.get(), n);
        } else if (op == 3) {
            rc = poms_dot(a0.get(), a1.get(), (int64_t)n, &dot, ws.get(), nullptr);
        } else if (op == 4) {   // z = a x + b y
            rc = poms_axpby(a0.get(), sc[0], a1.get(), sc[1], a2.get(), (int64_t)n, nullptr);
            put(a0.get(), n);
        } else if (op == 8) {   // y += sign * (num / den) * x
            rc = poms_axpy_dev(a0.get(), a1.get(), (int64_t)n, &sc[1], &sc[2], sc[0], nullptr);
            put(a0.get(), n);
        } else if (op == 9) {   // x = omega * b / d, fused sum of squares
            rc = poms_diag_scale(a0.get(), a1.get(), a2.get(), (int64_t)n, sc[0], has_dot ? &dot : nullptr, ws.get(), nullptr);
            put(a0.get(), n);
        } else if (op == 6) {   // d = c1 d + c2 z ; x += d
            rc = poms_cheb_update(a0.get(), a1.get(), a2.get(), sc[0], sc[3], (int64_t)n, nullptr);
            put(a0.get(), n);
            put(a1.get(), n);
        }
    }
    fclose(f);
    if (rc != 0) fprintf(stderr, "status %d: %s\n", rc, g_err);
    rc32 = rc;
    fseek(o, 0, SEEK_SET);
    fwrite(&rc32, 4, 1, o);
    fwrite(&dot, 8, 1, o);
    fclose(o);
    return 0;
}
