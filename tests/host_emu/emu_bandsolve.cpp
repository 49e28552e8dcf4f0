// emu_bandsolve.cpp -- TEST INFRASTRUCTURE ONLY (see cuda_emu.h): the banded line-solve kernels of
// poms_b200/csrc/poms_kernels.cu (section K4, rewritten for g++ by make_emu_source.py into
// band_solve_emu.cuh) run on the host through their C entry points.
//   emu_bandsolve <in> <out>
// in:  int32 header {variant (0 poms_band_solve_axis, 1 _fused, 2 _chunked), n, kl, ku, n_outer, s_outer,
//      s_axis, n_inner, has_piv, total, chunk, warm_fwd, warm_bwd, has_add, 0, 0}, fp64 scale,
//      ab ((2kl+ku+1) * n fp64), ipiv (n int32, if has_piv), y (total fp64), add (total fp64, if has_add)
// out: int32 status, x (total fp64)
#define POMS_HOST_EMU 1
#include "cuda_emu.h"
#include "band_solve_emu.cuh"

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
    const int variant = h[0], n = h[1], kl = h[2], ku = h[3];
    const int64_t n_outer = h[4], s_outer = h[5], s_axis = h[6], n_inner = h[7];
    const int has_piv = h[8], chunk = h[10], warm_f = h[11], warm_b = h[12], has_add = h[13];
    const size_t total = (size_t)h[9];
    auto sc = rd<double>(f, 1);
    auto ab = rd<double>(f, (size_t)(2 * kl + ku + 1) * n);
    auto piv = rd<int32_t>(f, has_piv ? n : 0);
    auto y = rd<double>(f, total);
    auto add = rd<double>(f, has_add ? total : 0);
    fclose(f);
    std::unique_ptr<double[]> x(new double[total]), work(new double[total]);
    for (size_t i = 0; i < total; ++i) x[i] = work[i] = 0.0;
    int rc;
    if (variant == 0)
        rc = poms_band_solve_axis(y.get(), x.get(), ab.get(), has_piv ? piv.get() : nullptr, n, kl, ku, n_outer,
                                  s_outer, s_axis, n_inner, nullptr);
    else if (variant == 1)
        rc = poms_band_solve_axis_fused(y.get(), work.get(), ab.get(), n, kl, ku, n_outer, s_outer, sc[0],
                                        has_add ? add.get() : nullptr, x.get(), nullptr);
    else
        rc = poms_band_solve_axis_chunked(y.get(), x.get(), work.get(), ab.get(), n, kl, ku, n_outer, s_outer,
                                          s_axis, n_inner, chunk, warm_f, warm_b, nullptr);
    if (rc != 0) fprintf(stderr, "status %d: %s\n", rc, g_err);
    FILE* o = fopen(argv[2], "wb");
    const int32_t rc32 = rc;
    fwrite(&rc32, 4, 1, o);
    fwrite(x.get(), 8, total, o);
    fclose(o);
    return 0;
}
