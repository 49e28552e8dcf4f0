// emu_transfer.cpp -- TEST INFRASTRUCTURE ONLY (see cuda_emu.h): runs poms_restrict_3d_v2 /
// poms_prolong_3d_v2 of poms_b200/csrc/poms_transfer3d_v2.cu, compiled for the host, on a problem read
// from a file.  Every array is copied into an exactly sized heap block, so that AddressSanitizer sees
// any access outside it.
//   emu_transfer <in> <out>
// in:  int32 header {op (0 restrict, 1 prolong: poms_*_3d_v2; 2, 3: the round-1 kernels poms_restrict_3d /
//      poms_prolong_3d of poms_transfer3d.cuh), n1f, n2f, n3f, n1c, n2c, n3c, W1, W2, W3, accumulate,
//      ldf, ldc, rows1, rows2, rows3}, then per axis starts (rows int32) and coefficients (rows * W
//      fp64), then the source array and the destination array (pitched, fp64).
// out: int32 status, then the destination array.
#define POMS_HOST_EMU 1
#include "../../poms_b200/csrc/poms_transfer3d_v2.cu"
#include "../../poms_b200/csrc/poms_transfer3d.cuh"

#include <cstdlib>
#include <memory>

template <class T>
static std::unique_ptr<T[]> rd(FILE* f, size_t n) {
    std::unique_ptr<T[]> p(new T[n]);
    if (fread(p.get(), sizeof(T), n, f) != n) {
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
    const int opx = h[0], op = opx & 1, n1f = h[1], n2f = h[2], n3f = h[3], n1c = h[4], n2c = h[5], n3c = h[6];
    const int W[3] = {h[7], h[8], h[9]}, accumulate = h[10], ldf = h[11], ldc = h[12];
    const int rows[3] = {h[13], h[14], h[15]};
    std::unique_ptr<int32_t[]> s[3];
    std::unique_ptr<double[]> c[3];
    for (int a = 0; a < 3; ++a) {
        s[a] = rd<int32_t>(f, rows[a]);
        c[a] = rd<double>(f, (size_t)rows[a] * W[a]);
    }
    const size_t nfine = (size_t)n1f * n2f * ldf, ncoarse = (size_t)n1c * n2c * ldc;
    auto src = rd<double>(f, op == 0 ? nfine : ncoarse);
    auto dst = rd<double>(f, op == 0 ? ncoarse : nfine);
    fclose(f);
    int rc;
    if (opx == 2)
        rc = poms_restrict_3d(src.get(), dst.get(), n1f, n2f, n3f, ldf, (int64_t)n2f * ldf, n1c, n2c, n3c, ldc,
                              (int64_t)n2c * ldc, s[0].get(), c[0].get(), W[0], s[1].get(), c[1].get(), W[1],
                              s[2].get(), c[2].get(), W[2], s[0].get(), s[1].get(), s[2].get(), nullptr);
    else if (opx == 3)
        rc = poms_prolong_3d(src.get(), dst.get(), n1f, n2f, n3f, ldf, (int64_t)n2f * ldf, n1c, n2c, n3c, ldc,
                             (int64_t)n2c * ldc, s[0].get(), c[0].get(), W[0], s[1].get(), c[1].get(), W[1],
                             s[2].get(), c[2].get(), W[2], s[1].get(), s[2].get(), accumulate, nullptr);
    else if (op == 0)
        rc = poms_restrict_3d_v2(src.get(), dst.get(), n1f, n2f, n3f, ldf, (int64_t)n2f * ldf, n1c, n2c, n3c,
                                 ldc, (int64_t)n2c * ldc, s[0].get(), c[0].get(), W[0], s[1].get(),
                                 c[1].get(), W[1], s[2].get(), c[2].get(), W[2], s[0].get(), s[1].get(),
                                 s[2].get(), nullptr);
    else
        rc = poms_prolong_3d_v2(src.get(), dst.get(), n1f, n2f, n3f, ldf, (int64_t)n2f * ldf, n1c, n2c, n3c,
                                ldc, (int64_t)n2c * ldc, s[0].get(), c[0].get(), W[0], s[1].get(),
                                c[1].get(), W[1], s[2].get(), c[2].get(), W[2], s[1].get(), s[2].get(),
                                accumulate, nullptr);
    if (rc != 0) fprintf(stderr, "status %d: %s\n", rc, g_err);
    FILE* o = fopen(argv[2], "wb");
    const int32_t rc32 = rc;
    fwrite(&rc32, 4, 1, o);
    fwrite(dst.get(), 8, op == 0 ? ncoarse : nfine, o);
    fclose(o);
    return 0;
}
