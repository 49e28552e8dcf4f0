// emu_dmma.cpp -- TEST INFRASTRUCTURE ONLY (see cuda_emu.h): poms_axis_dense_dmma (poms_extra.cu), the dense
// per-axis contraction of the coarse solve, with mma.sync.m8n8k4.f64 emulated lane by lane.
//   emu_dmma <in> <out>
// in:  int32 header (16): {n_in, n_out, 0...}, int64 (6): {n_outer, so_in, sa_in, so_out, sa_out, n_inner},
//      int64 (2): {len_in, len_out}, Q (n_out * n_in fp64), in (len_in fp64)
// out: int32 status, out (len_out fp64)
#define POMS_HOST_EMU 1
#include "cuda_emu.h"
#include "poms_b200.h"
static int x_bad_arg(int idx, const char* what) {
    snprintf(g_err, sizeof(g_err), "bad argument %d: %s", idx, what);
    return -idx;
}
static int x_fail_cuda(cudaError_t e, const char* where) {
    snprintf(g_err, sizeof(g_err), "%s: error %d", where, e);
    return e;
}
#include "dmma_emu.cuh"

#include <cstdlib>

int main(int argc, char** argv) {
    if (argc != 3) return 2;
    FILE* f = fopen(argv[1], "rb");
    if (!f) return 2;
    int32_t h[16];
    int64_t g[6], len[2];
    if (fread(h, 4, 16, f) != 16 || fread(g, 8, 6, f) != 6 || fread(len, 8, 2, f) != 2) return 3;
    const int n_in = h[0], n_out = h[1];
    double* Q = new double[(size_t)n_in * n_out];
    double* in = new double[len[0]];
    double* out = new double[len[1]];
    if (fread(Q, 8, (size_t)n_in * n_out, f) != (size_t)n_in * n_out || fread(in, 8, len[0], f) != (size_t)len[0]) return 3;
    fclose(f);
    for (int64_t i = 0; i < len[1]; ++i) out[i] = 0.0;
    const int rc = poms_axis_dense_dmma(in, out, Q, n_in, n_out, g[0], g[1], g[2], g[3], g[4], g[5], nullptr);
    if (rc != 0) fprintf(stderr, "status %d: %s\n", rc, g_err);
    FILE* o = fopen(argv[2], "wb");
    const int32_t rc32 = rc;
    fwrite(&rc32, 4, 1, o);
    fwrite(out, 8, len[1], o);
    fclose(o);
    delete[] Q;
    delete[] in;
    delete[] out;
    return 0;
}
