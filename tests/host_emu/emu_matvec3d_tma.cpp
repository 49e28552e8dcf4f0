// emu_matvec3d_tma.cpp -- TEST INFRASTRUCTURE ONLY (see cuda_emu.h, emu_tma.h): the hot path.  The
// TMA-staged 3-D Kronecker mat-vec kernels (kron_matvec3d_v3_kernel and the round-1
// kron_matvec3d_tma_kernel) run on the host through the C entry point poms_kron_matvec_3d_dotv -- tensor-map
// creation, Toeplitz tables, variant selection and chunking are the product's own host code.
//   emu_matvec3d_tma <in> <out>
// in:  int32 header (16): {p, form, epi, n1, n2, n3, ld, variant, has_b, has_dot, has_toep, force_generic,
//      dot_with, glo, ghi, 0} (x, and z, then hold glo + n1 + ghi planes: ghost planes of a slab), fp64 omega, bands m1 k1 m2 k2 m3 k3, toep coefficients (3*2*(2p+1) fp64) and
//      ranges (6 int32) if has_toep, x (n1*n2*ld), b (if has_b), z (if dot_with)
// out: int32 status, int32 fused flag, fp64 dot, fp64 number of emulated TMA copies, y
#define POMS_HOST_EMU 1
#include "cuda_emu.h"
#include "emu_tma.h"
#include "poms_b200.h"
#define POMS_WS_HEADER 256
#define POMS_MAX_PARTIALS 65536
#include "mv3_tma_emu.cuh"

#include <cstdlib>
#include <memory>

// 16-byte aligned, exactly sized blocks (the TMA path requires the alignment; ASan checks the size)
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
    const int p = h[0], form = h[1], epi = h[2], n1 = h[3], n2 = h[4], n3 = h[5], ld = h[6], variant = h[7];
    const int has_b = h[8], has_dot = h[9], has_toep = h[10], force_generic = h[11], dot_with = h[12], W = 2 * p + 1;
    const int glo = h[13], ghi = h[14];
    double om;
    rdinto(f, &om, 8);
    const int na[3] = {n1, n2, n3};
    std::unique_ptr<Buf> m[3], k[3];
    for (int a = 0; a < 3; ++a) {
        m[a].reset(new Buf((size_t)na[a] * W));
        rdinto(f, m[a]->p, (size_t)na[a] * W * 8);
        k[a].reset(new Buf((size_t)na[a] * W));
        rdinto(f, k[a]->p, (size_t)na[a] * W * 8);
    }
    double toep[3 * 2 * 11];
    int32_t rng[6];
    if (has_toep) {
        rdinto(f, toep, (size_t)3 * 2 * W * 8);
        rdinto(f, rng, sizeof(rng));
    }
    const size_t total = (size_t)n1 * n2 * ld, pl = (size_t)n2 * ld, totg = (size_t)(n1 + glo + ghi) * pl;
    Buf x(totg), b(has_b ? total : 0), z(dot_with ? totg : 0), y(total);
    rdinto(f, x.p, totg * 8);
    if (has_b) rdinto(f, b.p, total * 8);
    if (dot_with) rdinto(f, z.p, totg * 8);
    fclose(f);
    for (size_t i = 0; i < total; ++i) y.p[i] = 0.0;
    const size_t wsn = POMS_WS_HEADER + (size_t)POMS_MAX_PARTIALS * 8;
    std::unique_ptr<unsigned char[]> ws(new unsigned char[wsn]);
    memset(ws.get(), 0, wsn);
    double dot = 0.0;
    int fused = 0;
    poms_set_matvec3d_variant(variant);
    poms_set_force_generic(force_generic);
    const int rc = poms_kron_matvec_3d_dotv(x.p + glo * pl, y.p, has_b ? b.p : nullptr, n1, n2, n3, ld, (int64_t)pl, glo, ghi, p,
                                            form, m[0]->p, k[0]->p, m[1]->p, k[1]->p, m[2]->p, k[2]->p, epi, om,
                                            has_dot ? &dot : nullptr, ws.get(), nullptr, has_toep ? toep : nullptr,
                                            has_toep ? rng : nullptr, dot_with ? z.p + glo * pl : nullptr, &fused);
    if (rc != 0) fprintf(stderr, "status %d: %s\n", rc, g_err);
    FILE* o = fopen(argv[2], "wb");
    const int32_t rc32 = rc, fu = fused;
    fwrite(&rc32, 4, 1, o);
    fwrite(&fu, 4, 1, o);
    fwrite(&dot, 8, 1, o);
    const double nt = (double)emu_tma_loads;
    fwrite(&nt, 8, 1, o);
    fwrite(y.p, 8, total, o);
    fclose(o);
    return 0;
}
