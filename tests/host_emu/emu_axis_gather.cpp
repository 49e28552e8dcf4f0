// emu_axis_gather.cpp -- TEST INFRASTRUCTURE ONLY (see cuda_emu.h): poms_axis_gather (section K5 of
// poms_kernels.cu, rewritten for g++ by make_emu_source.py into axis_gather_emu.cuh): the per-axis sparse
// row gather behind the 2-D transfers, the fallback of the 3-D ones and the slab-plan rows.
//   emu_axis_gather <in> <out>
// in:  int32 header (16): {W, n_in, n_out, accumulate, rows, 0...}, int64 (6): {n_outer, so_in, sa_in, so_out,
//      sa_out, n_inner}, int64 (2): {len_in, len_out}, start (rows int32), coef (rows * W fp64),
//      in (len_in fp64), out (len_out fp64: the initial content, used when accumulate != 0)
// out: int32 status, out (len_out fp64)
#define POMS_HOST_EMU 1
#include "cuda_emu.h"
#include "poms_b200.h"
#include "axis_gather_emu.cuh"

#include <cstdlib>

int main(int argc, char** argv) {
    if (argc != 3) return 2;
    FILE* f = fopen(argv[1], "rb");
    if (!f) return 2;
    int32_t h[16];
    int64_t g[6], len[2];
    if (fread(h, 4, 16, f) != 16 || fread(g, 8, 6, f) != 6 || fread(len, 8, 2, f) != 2) return 3;
    const int W = h[0], n_in = h[1], n_out = h[2], acc = h[3], rows = h[4];
    int32_t* start = new int32_t[rows];
    double* coef = new double[(size_t)rows * W];
    double *in, *out;                                    // 16-byte aligned, exactly sized
    if (posix_memalign((void**)&in, 16, (size_t)len[0] * 8) || posix_memalign((void**)&out, 16, (size_t)len[1] * 8)) return 3;
    if (fread(start, 4, rows, f) != (size_t)rows || fread(coef, 8, (size_t)rows * W, f) != (size_t)rows * W ||
        fread(in, 8, len[0], f) != (size_t)len[0] || fread(out, 8, len[1], f) != (size_t)len[1])
        return 3;
    fclose(f);
    const int rc = poms_axis_gather(in, out, start, coef, W, n_in, n_out, g[0], g[1], g[2], g[3], g[4], g[5], acc, nullptr);
    if (rc != 0) fprintf(stderr, "status %d: %s\n", rc, g_err);
    FILE* o = fopen(argv[2], "wb");
    const int32_t rc32 = rc;
    fwrite(&rc32, 4, 1, o);
    fwrite(out, 8, len[1], o);
    fclose(o);
    delete[] start;
    delete[] coef;
    free(in);
    free(out);
    return 0;
}
