// emu_matvec3d.cpp -- TEST INFRASTRUCTURE ONLY (see cuda_emu.h): the generic 3-D Kronecker mat-vec
// (kron_matvec3d_kernel, translation unit 6 of poms_kernels.cu: the fallback for tiny or misaligned
// grids, with the deterministic grid reduction of the fused dot), rewritten for g++ by
// make_emu_source.py into mv3_generic_emu.cuh and launched through poms_mv3_generic_launch.
//   emu_matvec3d <in> <out>
// in:  int32 header {p, form, epi, n1, n2, n3, ld, chunk, has_b, has_dot, 0...} (16), fp64 omega,
//      bands m1 k1 m2 k2 m3 k3 (n_a * (2p+1) fp64 each), x (n1 * n2 * ld), b (same, if has_b)
// out: int32 status, fp64 dot, y (n1 * n2 * ld)
#define POMS_HOST_EMU 1
#include "cuda_emu.h"
#include "poms_b200.h"
#define POMS_WS_HEADER 256
#define POMS_MAX_PARTIALS 65536
#include "mv3_generic_emu.cuh"

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
    const int p = h[0], form = h[1], epi = h[2], n1 = h[3], n2 = h[4], n3 = h[5], ld = h[6], chunk = h[7];
    const int has_b = h[8], has_dot = h[9], W = 2 * p + 1;
    auto om = rd<double>(f, 1);
    const int na[3] = {n1, n2, n3};
    std::unique_ptr<double[]> m[3], k[3];
    for (int a = 0; a < 3; ++a) {
        m[a] = rd<double>(f, (size_t)na[a] * W);
        k[a] = rd<double>(f, (size_t)na[a] * W);
    }
    const size_t total = (size_t)n1 * n2 * ld;
    auto x = rd<double>(f, total);
    auto b = rd<double>(f, has_b ? total : 0);
    fclose(f);
    std::unique_ptr<double[]> y(new double[total]);
    for (size_t i = 0; i < total; ++i) y[i] = 0.0;
    const size_t wsn = POMS_WS_HEADER + (size_t)POMS_MAX_PARTIALS * 8;
    std::unique_ptr<unsigned char[]> ws(new unsigned char[wsn]);
    memset(ws.get(), 0, wsn);
    double dot = 0.0;
    MV3 a;
    a.x = x.get(); a.y = y.get(); a.b = has_b ? b.get() : nullptr;
    a.n1 = n1; a.n2 = n2; a.n3 = n3; a.ld = ld; a.pld = (int64_t)n2 * ld; a.glo = a.ghi = 0;
    a.m1 = m[0].get(); a.k1 = k[0].get(); a.m2 = m[1].get(); a.k2 = k[1].get(); a.m3 = m[2].get(); a.k3 = k[2].get();
    a.omega = om[0]; a.dot_out = has_dot ? &dot : nullptr; a.ws = ws.get(); a.chunk = chunk;
    dim3 grid((n3 + 63) / 64, (n2 + 15) / 16, (n1 + chunk - 1) / chunk);
    const int rc = poms_mv3_generic_launch(a, p, form, epi, grid, nullptr);
    if (rc != 0) fprintf(stderr, "status %d: %s\n", rc, g_err);
    FILE* o = fopen(argv[2], "wb");
    const int32_t rc32 = rc;
    fwrite(&rc32, 4, 1, o);
    fwrite(&dot, 8, 1, o);
    fwrite(y.get(), 8, total, o);
    fclose(o);
    return 0;
}
