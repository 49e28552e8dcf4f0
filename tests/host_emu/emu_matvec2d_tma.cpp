// emu_matvec2d_tma.cpp -- TEST INFRASTRUCTURE ONLY (see cuda_emu.h, emu_tma.h): the 2-D fast path.  The
// warp-autonomous TMA mat-vec (kron_matvec2d_tma_kernel, translation unit 7) runs on the host through
// poms_kron_matvec_2d_ex; tensor map, Toeplitz tables and chunking are the product's own host code.
//   emu_matvec2d_tma <in> <out>
// in:  int32 header (16): {p, form, epi, n1, n2, ld, variant, has_b, has_dot, has_toep, 0...}, fp64 omega,
//      bands m1 k1 m2 k2, toep coefficients (2*2*(2p+1) fp64) and ranges (4 int32) if has_toep,
//      x (n1*ld), b (if has_b)
// out: int32 status, int32 0, fp64 dot, fp64 number of emulated TMA copies, y
#define POMS_HOST_EMU 1
#include "cuda_emu.h"
#include "emu_tma.h"
#include "poms_b200.h"
#define POMS_WS_HEADER 256
#define POMS_MAX_PARTIALS 65536
#include "mv2_tma_emu.cuh"

#include <cstdlib>
#include <memory>

struct Buf {
    double* p = nullptr;
    explicit Buf(size_t n) { if (posix_memalign((void**)&p, 128, (n ? n : 1) * 8)) abort(); }
    ~Buf() { free(p); }
};
static void rdinto(FILE* f, void* dst, size_t bytes) {
    if (bytes && fread(dst, 1, bytes, f) != bytes) {
        fprintf(stderr, "short read\n");
        exit(3);
    }
}

int main(int argc, char** argv) {
    if (argc != 3) return 2;
    FILE* f = fopen(argv[1], "rb");
    if (!f) return 2;
    int32_t h[16];
    rdinto(f, h, sizeof(h));
    const int p = h[0], form = h[1], epi = h[2], n1 = h[3], n2 = h[4], ld = h[5], variant = h[6], has_b = h[7],
              has_dot = h[8], has_toep = h[9], W = 2 * p + 1;
    double om;
    rdinto(f, &om, 8);
    Buf m1((size_t)n1 * W), k1((size_t)n1 * W), m2((size_t)n2 * W), k2((size_t)n2 * W);
    rdinto(f, m1.p, (size_t)n1 * W * 8);
    rdinto(f, k1.p, (size_t)n1 * W * 8);
    rdinto(f, m2.p, (size_t)n2 * W * 8);
    rdinto(f, k2.p, (size_t)n2 * W * 8);
    double toep[2 * 2 * 11];
    int32_t rng[4];
    if (has_toep) {
        rdinto(f, toep, (size_t)2 * 2 * W * 8);
        rdinto(f, rng, sizeof(rng));
    }
    const size_t total = (size_t)n1 * ld;
    Buf x(total), b(has_b ? total : 0), y(total);
    rdinto(f, x.p, total * 8);
    if (has_b) rdinto(f, b.p, total * 8);
    fclose(f);
    for (size_t i = 0; i < total; ++i) y.p[i] = 0.0;
    const size_t wsn = POMS_WS_HEADER + (size_t)POMS_MAX_PARTIALS * 8;
    std::unique_ptr<unsigned char[]> ws(new unsigned char[wsn]);
    memset(ws.get(), 0, wsn);
    double dot = 0.0;
    poms_set_matvec2d_variant(variant);
    const int rc = poms_kron_matvec_2d_ex(x.p, y.p, has_b ? b.p : nullptr, n1, n2, ld, 0, 0, p, form, m1.p, k1.p, m2.p,
                                          k2.p, epi, om, has_dot ? &dot : nullptr, ws.get(), nullptr,
                                          has_toep ? toep : nullptr, has_toep ? rng : nullptr);
    if (rc != 0) fprintf(stderr, "status %d: %s\n", rc, g_err);
    FILE* o = fopen(argv[2], "wb");
    const int32_t rc32 = rc, zero = 0;
    fwrite(&rc32, 4, 1, o);
    fwrite(&zero, 4, 1, o);
    fwrite(&dot, 8, 1, o);
    const double nt = (double)emu_tma_loads;
    fwrite(&nt, 8, 1, o);
    fwrite(y.p, 8, total, o);
    fclose(o);
    return 0;
}
