// emu_setup.cpp -- TEST INFRASTRUCTURE ONLY (see cuda_emu.h): the device-side 1-D set-up kernels of
// poms_setup.cu (rewritten for g++ by make_emu_source.py into setup_emu.cuh), exactly sized buffers.
//   emu_setup <in> <out>
// in:  int32 header (16): {op, n, p, nc, nf, q, 0...}, then K fp64 arrays, each as int64 length + data
//      op 0  poms_assemble_1d:          knots (n+p+1), gauss_x (p+1), gauss_w (p+1)   -> M, K (n * (2p+1) each)
//      op 1  poms_knot_insertion_rows:  Tc (nc+p+1), Tf (nf+p+1)                      -> start (nf int32), coef (nf * (p+1))
//      op 2  poms_band_lu_nopiv:        band (n * (2q+1))                             -> info (int32), ab ((3q+1) * n)
// out: int32 status, then the outputs in that order
#define POMS_HOST_EMU 1
#include "cuda_emu.h"
#include "poms_b200.h"
#include "setup_emu.cuh"

#include <cstdlib>
#include <vector>

static double* read_arr(FILE* f, int64_t expect) {
    int64_t len;
    if (fread(&len, 8, 1, f) != 1 || len != expect) exit(3);
    double* a = new double[len];                   // exactly sized: the sanitizer sees every overrun
    if (fread(a, 8, len, f) != (size_t)len) exit(3);
    return a;
}

int main(int argc, char** argv) {
    if (argc != 3) return 2;
    FILE* f = fopen(argv[1], "rb");
    if (!f) return 2;
    int32_t h[16];
    if (fread(h, 4, 16, f) != 16) return 3;
    const int op = h[0], n = h[1], p = h[2], nc = h[3], nf = h[4], q = h[5];
    FILE* o = fopen(argv[2], "wb");
    int32_t rc = -99;
    if (op == 0) {
        double* T = read_arr(f, n + p + 1);
        double* gx = read_arr(f, p + 1);
        double* gw = read_arr(f, p + 1);
        const size_t len = (size_t)n * (2 * p + 1);
        double *M = new double[len], *K = new double[len];
        rc = poms_assemble_1d(T, n, p, gx, gw, M, K, nullptr);
        fwrite(&rc, 4, 1, o);
        fwrite(M, 8, len, o);
        fwrite(K, 8, len, o);
        delete[] T; delete[] gx; delete[] gw; delete[] M; delete[] K;
    } else if (op == 1) {
        double* Tc = read_arr(f, nc + p + 1);
        double* Tf = read_arr(f, nf + p + 1);
        int32_t* start = new int32_t[nf];
        double* coef = new double[(size_t)nf * (p + 1)];
        rc = poms_knot_insertion_rows(Tc, nc, Tf, nf, p, start, coef, nullptr);
        fwrite(&rc, 4, 1, o);
        fwrite(start, 4, nf, o);
        fwrite(coef, 8, (size_t)nf * (p + 1), o);
        delete[] Tc; delete[] Tf; delete[] start; delete[] coef;
    } else if (op == 2) {
        double* band = read_arr(f, (int64_t)n * (2 * q + 1));
        double* ab = new double[(size_t)(3 * q + 1) * n];
        int* info = new int[1];
        info[0] = -7;
        rc = poms_band_lu_nopiv(band, n, q, ab, info, nullptr);
        fwrite(&rc, 4, 1, o);
        fwrite(info, 4, 1, o);
        fwrite(ab, 8, (size_t)(3 * q + 1) * n, o);
        delete[] band; delete[] ab; delete[] info;
    }
    if (rc != 0) fprintf(stderr, "status %d: %s\n", rc, g_err);
    fclose(f);
    fclose(o);
    return 0;
}
