// poms_kernels.cu -- hand-written sm_100a fp64 kernels + C ABI for the POMS multigrid solve path.
// See include/poms_b200.h for the contract and the reference call sites each entry replaces.
//
// Design notes (DESIGN.md has the long form):
//  * The Kronecker mat-vec is ONE pass over HBM: a CTA owns a (rows x cols) tile of the fast
//    axes and marches along axis 1 (slowest).  The in-row band pass goes through shared
//    memory, the cross-row pass through a small shared window, and the axis-1 pass is kept in
//    REGISTERS as 2p+1 rotating partial sums per owned point (scatter form), so every x is
//    read once and every y written once (16 B/DOF).
//  * Reductions are deterministic: per-CTA partials in a fixed slot + last-CTA tree sum.
//  * Everything is HBM-bound integer-free fp64 streaming; no tensor cores here by design.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include "poms_b200.h"

#define POMS_MAX_PARTIALS 65536
#define POMS_WS_HEADER 256

// The library is built from several translation units of THIS file (build.py: -DPOMS_TU=k) so that
// the heavy template families compile in parallel: 0 = everything else, 1..5 = TMA 3-D mat-vec of
// degree k, 6 = generic 3-D mat-vec.  Error text and launch counter live in TU 0.
#ifndef POMS_TU
#define POMS_TU 0
#endif
#define POMS_HIDDEN __attribute__((visibility("hidden")))
#if POMS_TU == 0
POMS_HIDDEN thread_local char g_err[256] = "";
POMS_HIDDEN int64_t g_launches = 0;
#else
extern POMS_HIDDEN thread_local char g_err[256];
extern POMS_HIDDEN int64_t g_launches;
#endif

[[maybe_unused]] static int fail_cuda(cudaError_t e, const char* where) {
    snprintf(g_err, sizeof(g_err), "%s: %s", where, cudaGetErrorString(e));
    return (int)e;
}
[[maybe_unused]] static int bad_arg(int idx, const char* what) {
    snprintf(g_err, sizeof(g_err), "bad argument %d: %s", idx, what);
    return -idx;
}
#define CHECK_LAUNCH(where)                                     \
    do {                                                        \
        g_launches++;                                           \
        cudaError_t e_ = cudaGetLastError();                    \
        if (e_ != cudaSuccess) return fail_cuda(e_, where);     \
    } while (0)
// launch of a 256-thread kernel with one struct argument; tests/host_emu redefines it to run the
// kernel source on the host under sanitizers (POMS_HOST_EMU, never defined by build.py)
#define POMS_LAUNCH(kernel, grid, stream, arg) kernel<<<grid, 256, 0, stream>>>(arg)

#if POMS_TU == 0
extern "C" int poms_version(void) { return 100; }
extern "C" int64_t poms_workspace_bytes(void) {
    return POMS_WS_HEADER + (int64_t)POMS_MAX_PARTIALS * sizeof(double);
}
extern "C" const char* poms_last_error(void) { return g_err; }
extern "C" int64_t poms_launch_count(void) { return g_launches; }
extern "C" void poms_launch_count_add(int64_t n) { g_launches += n; }
#endif

// ------------------------------------------------------------------------------------------
// cp.async helpers (Ampere-style asynchronous global -> shared copies, 8 bytes per thread)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async8(void* smem, const void* gmem) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(sa), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

// ------------------------------------------------------------------------------------------
// deterministic grid reduction
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Sum `v` over the CTA (result valid in thread 0).  `red` holds >= 32 doubles.
__device__ __forceinline__ double block_sum(double v, double* red) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) red[w] = v;
    __syncthreads();
    double s = 0.0;
    if (w == 0) {
        s = lane < nw ? red[lane] : 0.0;
        s = warp_sum(s);
    }
    return s;
}

// Every CTA deposits its partial; the last one to arrive sums all partials in a fixed order
// and writes *out.  ws: [0] ticket counter (uint), partials from byte POMS_WS_HEADER.
__device__ __forceinline__ void grid_sum_finish(double block_total, double* out, void* ws,
                                                unsigned nblocks, unsigned bid, double* red) {
    unsigned* ticket = (unsigned*)ws;
    double* part = (double*)((char*)ws + POMS_WS_HEADER);
    __shared__ unsigned s_last;
    if (threadIdx.x == 0) {
        part[bid] = block_total;
        __threadfence();
        unsigned t = atomicAdd(ticket, 1u);
        s_last = (t == nblocks - 1) ? 1u : 0u;
    }
    __syncthreads();
    if (s_last) {
        __threadfence();
        double s = 0.0;
        for (unsigned i = threadIdx.x; i < nblocks; i += blockDim.x) s += __ldcg(part + i);
        s = block_sum(s, red);
        if (threadIdx.x == 0) {
            *out = s;
            *ticket = 0u;
        }
    }
}

// Rotating partial sums of the marching-axis band pass.  Input line t (slot u = t % W) adds
// c[k]*value into the partial sum of output line t + P - k, which lives in slot (u+W-1-k) % W;
// the output completed by this line is slot u: it is returned in vout and cleared.  The switch
// keeps every register index static.
template <int W, int E, bool TWO, int U>
__device__ __forceinline__ void rot_scatter_u(double (&acc)[E][W], const double (&ta)[E],
                                              const double (&tb)[E], const double (&c1k)[W],
                                              const double (&c1m)[W], double (&vout)[E]) {
#pragma unroll
    for (int e = 0; e < E; ++e) {
#pragma unroll
        for (int k = 0; k < W; ++k) {
            constexpr int dummy = 0;
            (void)dummy;
            const int slot = (U + W - 1 - k) % W;
            acc[e][slot] = fma(c1k[k], ta[e], acc[e][slot]);
            if (TWO) acc[e][slot] = fma(c1m[k], tb[e], acc[e][slot]);
        }
        vout[e] = acc[e][U];
        acc[e][U] = 0.0;
    }
}
template <int W, int E, bool TWO>
__device__ __forceinline__ void rot_scatter(int u, double (&acc)[E][W], const double (&ta)[E],
                                            const double (&tb)[E], const double (&c1k)[W],
                                            const double (&c1m)[W], double (&vout)[E]) {
    switch (u) {
        case 0: rot_scatter_u<W, E, TWO, 0>(acc, ta, tb, c1k, c1m, vout); break;
        case 1: rot_scatter_u<W, E, TWO, 1>(acc, ta, tb, c1k, c1m, vout); break;
        case 2: rot_scatter_u<W, E, TWO, 2>(acc, ta, tb, c1k, c1m, vout); break;
        case 3: if (W > 3) rot_scatter_u<W, E, TWO, (3 < W ? 3 : 0)>(acc, ta, tb, c1k, c1m, vout); break;
        case 4: if (W > 4) rot_scatter_u<W, E, TWO, (4 < W ? 4 : 0)>(acc, ta, tb, c1k, c1m, vout); break;
        case 5: if (W > 5) rot_scatter_u<W, E, TWO, (5 < W ? 5 : 0)>(acc, ta, tb, c1k, c1m, vout); break;
        case 6: if (W > 6) rot_scatter_u<W, E, TWO, (6 < W ? 6 : 0)>(acc, ta, tb, c1k, c1m, vout); break;
        case 7: if (W > 7) rot_scatter_u<W, E, TWO, (7 < W ? 7 : 0)>(acc, ta, tb, c1k, c1m, vout); break;
        case 8: if (W > 8) rot_scatter_u<W, E, TWO, (8 < W ? 8 : 0)>(acc, ta, tb, c1k, c1m, vout); break;
        case 9: if (W > 9) rot_scatter_u<W, E, TWO, (9 < W ? 9 : 0)>(acc, ta, tb, c1k, c1m, vout); break;
        default: if (W > 10) rot_scatter_u<W, E, TWO, (10 < W ? 10 : 0)>(acc, ta, tb, c1k, c1m, vout); break;
    }
}

// Axis-1 partial sums in SHIFT form: after input plane j, acc[e][m] is the partial sum of output plane
// j - P + m.  Plane j adds c[W-1-m] * value to it; in place that reads acc[m+1] (the same output one
// plane earlier) and writes acc[m], for ascending m, so the window slides without a rotation index,
// without register moves and without the 7-way switch of rot_scatter: one straight-line block of
// W (2W for the sum form) independent FMA chains per point.  The completed plane is acc[e][0].
template <int W, int E, bool TWO>
__device__ __forceinline__ void shift_scatter(double (&acc)[E][W], const double (&ta)[E], const double (&tb)[E],
                                              const double (&c1k)[W], const double (&c1m)[W], double (&vout)[E]) {
#pragma unroll
    for (int e = 0; e < E; ++e) {
#pragma unroll
        for (int m = 0; m < W - 1; ++m) {
            acc[e][m] = fma(c1k[W - 1 - m], ta[e], acc[e][m + 1]);
            if (TWO) acc[e][m] = fma(c1m[W - 1 - m], tb[e], acc[e][m]);
        }
        acc[e][W - 1] = c1k[0] * ta[e];
        if (TWO) acc[e][W - 1] = fma(c1m[0], tb[e], acc[e][W - 1]);
        vout[e] = acc[e][0];
    }
}

struct MV2 {
    const double* x;
    double* y;
    const double* b;
    int n1, n2;
    int64_t ld;
    int glo, ghi;
    const double *m1, *k1, *m2, *k2;
    double omega;
    double* dot_out;
    void* ws;
    int chunk;
};

// ------------------------------------------------------------------------------------------
// K1: Kronecker banded mat-vec, 3-D
// ------------------------------------------------------------------------------------------
struct MV3 {
    const double* x;
    double* y;
    const double* b;
    int n1, n2, n3;
    int64_t ld, pld;
    int glo, ghi;
    const double *m1, *k1, *m2, *k2, *m3, *k3;
    double omega;
    double* dot_out;
    void* ws;
    int chunk;  // output planes per CTA along axis 1
};

POMS_HIDDEN int poms_mv3_generic_launch(const MV3& a, int p, int form, int epi, dim3 grid, cudaStream_t st);

#if POMS_TU == 6
template <int P, int FORM, int EPI>
__global__ void __launch_bounds__(256, 2) kron_matvec3d_kernel(MV3 a) {
    constexpr int W = 2 * P + 1;
    constexpr int T3 = 64, TY = 4, E = 4, T2 = TY * E;
    constexpr int R2 = T2 + 2 * P;  // staged rows
    constexpr int C3 = T3 + 2 * P;  // staged cols
    __shared__ double sx[R2][C3];
    __shared__ double su[R2][T3];
    __shared__ double sv[FORM == POMS_FORM_SUM ? R2 : 1][T3];
    __shared__ double c2m[T2][W];
    __shared__ double c2k[FORM == POMS_FORM_SUM ? T2 : 1][W];
    __shared__ double red[32];

    const int tx = threadIdx.x & (T3 - 1), ty = threadIdx.x / T3;
    const int i3_0 = blockIdx.x * T3, i2_0 = blockIdx.y * T2;
    const int c_lo = blockIdx.z * a.chunk;
    const int c_hi = min(a.n1, c_lo + a.chunk);
    const int i3 = i3_0 + tx;
    const bool v3 = i3 < a.n3;

    // axis-3 coefficients of this thread's column (registers)
    double m3c[W], k3c[W];
#pragma unroll
    for (int k = 0; k < W; ++k) {
        m3c[k] = v3 ? a.m3[(int64_t)i3 * W + k] : 0.0;
        k3c[k] = (FORM == POMS_FORM_SUM && v3) ? a.k3[(int64_t)i3 * W + k] : 0.0;
    }
    // axis-2 coefficients of the tile rows (shared)
    for (int t = threadIdx.x; t < T2 * W; t += blockDim.x) {
        const int r = t / W, k = t % W, i2 = i2_0 + r;
        c2m[r][k] = i2 < a.n2 ? a.m2[(int64_t)i2 * W + k] : 0.0;
        if (FORM == POMS_FORM_SUM) c2k[r][k] = i2 < a.n2 ? a.k2[(int64_t)i2 * W + k] : 0.0;
    }
    // diagonal pieces for the Jacobi epilogue
    double dA[E], dB[E];
    if (EPI == POMS_EPI_JACOBI || EPI == POMS_EPI_DINV) {
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const int i2 = i2_0 + ty * E + e;
            const bool ok = v3 && i2 < a.n2;
            const double m2d = ok ? a.m2[(int64_t)i2 * W + P] : 1.0;
            const double m3d = ok ? a.m3[(int64_t)i3 * W + P] : 1.0;
            if (FORM == POMS_FORM_SUM) {
                const double k2d = ok ? a.k2[(int64_t)i2 * W + P] : 0.0;
                const double k3d = ok ? a.k3[(int64_t)i3 * W + P] : 0.0;
                dA[e] = m2d * m3d;
                dB[e] = k2d * m3d + m2d * k3d;
            } else {
                dA[e] = m2d * m3d;
                dB[e] = 0.0;
            }
        }
    }

    double acc[E][W];
#pragma unroll
    for (int e = 0; e < E; ++e)
#pragma unroll
        for (int k = 0; k < W; ++k) acc[e][k] = 0.0;
    double dsum = 0.0;

    const int start = c_lo - P;     // first input plane
    const int end = c_hi + P;       // one past the last input plane
    int u = 0;                      // (j1 - start) % W: slot of the output this plane completes
    for (int j1 = start; j1 < end; ++j1) {
        const bool have = (j1 >= -a.glo) && (j1 < a.n1 + a.ghi);
        double ta[E], tb[E], c1k[W], c1m[W];
        if (have) {
            // ---- stage 0: stage the halo'd plane tile ----
            const double* xp = a.x + (int64_t)j1 * a.pld;
            for (int t = threadIdx.x; t < R2 * C3; t += blockDim.x) {
                const int r = t / C3, c = t - r * C3;
                const int g2 = i2_0 - P + r, g3 = i3_0 - P + c;
                double v = 0.0;
                if (g2 >= 0 && g2 < a.n2 && g3 >= 0 && g3 < a.n3)
                    v = __ldg(xp + (int64_t)g2 * a.ld + g3);
                sx[r][c] = v;
            }
            __syncthreads();
            // ---- stage 1: band pass along axis 3 ----
            for (int r = ty; r < R2; r += TY) {
                double uu = 0.0, vv = 0.0;
#pragma unroll
                for (int k = 0; k < W; ++k) {
                    const double xv = sx[r][tx + k];
                    uu = fma(m3c[k], xv, uu);
                    if (FORM == POMS_FORM_SUM) vv = fma(k3c[k], xv, vv);
                }
                su[r][tx] = uu;
                if (FORM == POMS_FORM_SUM) sv[r][tx] = vv;
            }
            __syncthreads();
            // ---- stage 2: band pass along axis 2 (window in registers) ----
            double wu[E + 2 * P], wv[E + 2 * P];
#pragma unroll
            for (int r = 0; r < E + 2 * P; ++r) {
                wu[r] = su[ty * E + r][tx];
                if (FORM == POMS_FORM_SUM) wv[r] = sv[ty * E + r][tx];
            }
#pragma unroll
            for (int e = 0; e < E; ++e) {
                const int r2 = ty * E + e;
                double sa = 0.0, sb = 0.0;
#pragma unroll
                for (int k = 0; k < W; ++k) {
                    const double cm = c2m[r2][k];
                    sa = fma(cm, wu[e + k], sa);
                    if (FORM == POMS_FORM_SUM) {
                        sb = fma(c2k[r2][k], wu[e + k], sb);
                        sb = fma(cm, wv[e + k], sb);
                    }
                }
                ta[e] = sa;
                tb[e] = sb;
            }
            // coefficients of the axis-1 pass for this input plane (block uniform)
#pragma unroll
            for (int k = 0; k < W; ++k) {
                const int i1 = j1 + P - k;  // output plane fed through band column k
                const bool ok = i1 >= 0 && i1 < a.n1;
                if (FORM == POMS_FORM_SUM) {
                    c1k[k] = ok ? __ldg(a.k1 + (int64_t)i1 * W + k) : 0.0;
                    c1m[k] = ok ? __ldg(a.m1 + (int64_t)i1 * W + k) : 0.0;
                } else {
                    c1k[k] = ok ? __ldg(a.m1 + (int64_t)i1 * W + k) : 0.0;
                    c1m[k] = 0.0;
                }
            }
        } else {
#pragma unroll
            for (int e = 0; e < E; ++e) ta[e] = tb[e] = 0.0;
#pragma unroll
            for (int k = 0; k < W; ++k) c1k[k] = c1m[k] = 0.0;
        }
        // ---- stage 3: scatter into the rotating axis-1 partial sums; slot u completes ----
        double vout[E];
        rot_scatter<W, E, FORM == POMS_FORM_SUM>(u, acc, ta, tb, c1k, c1m, vout);
        const int i1 = j1 - P;
        if (i1 >= c_lo && i1 < c_hi) {
#pragma unroll
            for (int e = 0; e < E; ++e) {
                const int i2 = i2_0 + ty * E + e;
                if (v3 && i2 < a.n2) {
                    const int64_t idx = (int64_t)i1 * a.pld + (int64_t)i2 * a.ld + i3;
                    const double v = vout[e];
                    if (EPI == POMS_EPI_STORE) {
                        a.y[idx] = v;
                        if (a.dot_out) dsum = fma(__ldg(a.x + idx), v, dsum);
                    } else if (EPI == POMS_EPI_RESID) {
                        const double rr = a.b[idx] - v;
                        a.y[idx] = rr;
                        dsum = fma(rr, rr, dsum);
                    } else if (EPI == POMS_EPI_AXPY) {
                        const double w_ = a.omega * v;
                        a.y[idx] = a.b ? a.b[idx] + w_ : w_;
                        dsum = fma(w_, w_, dsum);
                    } else {
                        double dg;
                        if (FORM == POMS_FORM_SUM)
                            dg = __ldg(a.k1 + (int64_t)i1 * W + P) * dA[e] +
                                 __ldg(a.m1 + (int64_t)i1 * W + P) * dB[e];
                        else
                            dg = __ldg(a.m1 + (int64_t)i1 * W + P) * dA[e];
                        const double dr = a.omega * (a.b[idx] - v) / dg;
                        a.y[idx] = (EPI == POMS_EPI_JACOBI) ? __ldg(a.x + idx) + dr : dr;
                        dsum = fma(dr, dr, dsum);
                    }
                }
            }
        }
        u = (u + 1 == W) ? 0 : u + 1;
    }
    if (a.dot_out) {
        const double tot = block_sum(dsum, red);
        const unsigned nb = gridDim.x * gridDim.y * gridDim.z;
        const unsigned bid = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
        grid_sum_finish(tot, a.dot_out, a.ws, nb, bid, red);
    }
}

template <int P, int FORM>
static int launch_mv3_epi(const MV3& a, int epi, dim3 grid, cudaStream_t st) {
    switch (epi) {
        case POMS_EPI_STORE: kron_matvec3d_kernel<P, FORM, POMS_EPI_STORE><<<grid, 256, 0, st>>>(a); break;
        case POMS_EPI_RESID: kron_matvec3d_kernel<P, FORM, POMS_EPI_RESID><<<grid, 256, 0, st>>>(a); break;
        case POMS_EPI_JACOBI: kron_matvec3d_kernel<P, FORM, POMS_EPI_JACOBI><<<grid, 256, 0, st>>>(a); break;
        case POMS_EPI_DINV: kron_matvec3d_kernel<P, FORM, POMS_EPI_DINV><<<grid, 256, 0, st>>>(a); break;
        case POMS_EPI_AXPY: kron_matvec3d_kernel<P, FORM, POMS_EPI_AXPY><<<grid, 256, 0, st>>>(a); break;
        default: return bad_arg(19, "epilogue");
    }
    return 0;
}
template <int P>
static int launch_mv3(const MV3& a, int form, int epi, dim3 grid, cudaStream_t st) {
    if (form == POMS_FORM_SINGLE) return launch_mv3_epi<P, POMS_FORM_SINGLE>(a, epi, grid, st);
    if (form == POMS_FORM_SUM) return launch_mv3_epi<P, POMS_FORM_SUM>(a, epi, grid, st);
    return bad_arg(12, "form");
}
int poms_mv3_generic_launch(const MV3& a, int p, int form, int epi, dim3 grid, cudaStream_t st) {
    switch (p) {
        case 1: return launch_mv3<1>(a, form, epi, grid, st);
        case 2: return launch_mv3<2>(a, form, epi, grid, st);
        case 3: return launch_mv3<3>(a, form, epi, grid, st);
        case 4: return launch_mv3<4>(a, form, epi, grid, st);
        case 5: return launch_mv3<5>(a, form, epi, grid, st);
        default: return bad_arg(11, "p must be 1..5");
    }
}
#endif  // POMS_TU == 6

#if POMS_TU == 0
static int g_chunk_override = 0;   // > 0: fixed axis-1 chunk (A/B timing only)
extern "C" void poms_set_matvec3d_chunk(int c) { g_chunk_override = c; }
static int pick_chunk(int n1, int64_t tiles, int p) {
    if (g_chunk_override > 0) return g_chunk_override < n1 ? g_chunk_override : n1;
    // enough CTAs for ~8 waves of 148 SMs x 2 CTAs on big grids (boundary and ragged tiles are cheaper
    // than interior ones, so short CTAs balance better: 65 planes measured 3 % faster than 129 at
    // 515^3), ~4 waves on small ones, but keep the 2p halo planes amortised:
    // big grids <= 12.5 % redundant halo planes; coarse-level grids are latency bound, so more,
    // shorter CTAs win even at 50-100 % redundancy (measured sweep: profiles/r01_ab_chunk_sweep.txt)
    const int64_t want = 148 * 2 * (tiles >= 148 ? 8 : 4);
    int64_t nch = (want + tiles - 1) / tiles;
    if (nch < 1) nch = 1;
    int chunk = (int)((n1 + nch - 1) / nch);
    const int min_chunk = (tiles >= 148 ? 8 : tiles >= 20 ? 2 : 1) * (2 * p);
    if (chunk < min_chunk) chunk = min_chunk;
    if (chunk > n1) chunk = n1;
    return chunk;
}
#endif

#include "poms_matvec3d_tma.cuh"

#if POMS_TU == 0
// axis-1 chunk of the 2-D TMA kernel: ~2 waves of 148 SMs x 5 CTAs, at least 12p rows per chunk
static int pick_chunk2d(int n1, int64_t strips, int p) {
    int64_t nch = (148 * 5 * 2 + strips - 1) / strips;
    if (nch < 1) nch = 1;
    int chunk = (int)((n1 + nch - 1) / nch);
    if (chunk < 12 * p) chunk = 12 * p;
    if (chunk > n1) chunk = n1;
    return chunk;
}
#endif
#include "poms_matvec2d_tma.cuh"

#if POMS_TU == 0

extern "C" int poms_kron_matvec_3d_ex(const double* x, double* y, const double* b, int n1, int n2,
                                      int n3, int64_t ld, int64_t pld, int glo, int ghi, int p,
                                      int form, const double* m1, const double* k1,
                                      const double* m2, const double* k2, const double* m3,
                                      const double* k3, int epilogue, double omega,
                                      double* dot_out, void* ws, void* stream,
                                      const double* toep_host, const int* toep_rng_host);

extern "C" int poms_kron_matvec_3d(const double* x, double* y, const double* b, int n1, int n2,
                                   int n3, int64_t ld, int64_t pld, int glo, int ghi, int p,
                                   int form, const double* m1, const double* k1,
                                   const double* m2, const double* k2, const double* m3,
                                   const double* k3, int epilogue, double omega,
                                   double* dot_out, void* ws, void* stream) {
    return poms_kron_matvec_3d_ex(x, y, b, n1, n2, n3, ld, pld, glo, ghi, p, form, m1, k1, m2, k2,
                                  m3, k3, epilogue, omega, dot_out, ws, stream, nullptr, nullptr);
}

static int g_force_generic = 0;
extern "C" void poms_set_force_generic(int flag) { g_force_generic = flag; }

extern "C" int poms_kron_matvec_3d_dotv(const double* x, double* y, const double* b, int n1, int n2,
                                        int n3, int64_t ld, int64_t pld, int glo, int ghi, int p,
                                        int form, const double* m1, const double* k1,
                                        const double* m2, const double* k2, const double* m3,
                                        const double* k3, int epilogue, double omega,
                                        double* dot_out, void* ws, void* stream,
                                        const double* toep_host, const int* toep_rng_host,
                                        const double* dot_with, int* fused_host);

extern "C" int poms_kron_matvec_3d_ex(const double* x, double* y, const double* b, int n1, int n2,
                                      int n3, int64_t ld, int64_t pld, int glo, int ghi, int p,
                                      int form, const double* m1, const double* k1,
                                      const double* m2, const double* k2, const double* m3,
                                      const double* k3, int epilogue, double omega,
                                      double* dot_out, void* ws, void* stream,
                                      const double* toep_host, const int* toep_rng_host) {
    return poms_kron_matvec_3d_dotv(x, y, b, n1, n2, n3, ld, pld, glo, ghi, p, form, m1, k1, m2, k2, m3, k3,
                                    epilogue, omega, dot_out, ws, stream, toep_host, toep_rng_host, nullptr,
                                    nullptr);
}

extern "C" int poms_kron_matvec_3d_dotv(const double* x, double* y, const double* b, int n1, int n2,
                                        int n3, int64_t ld, int64_t pld, int glo, int ghi, int p,
                                        int form, const double* m1, const double* k1,
                                        const double* m2, const double* k2, const double* m3,
                                        const double* k3, int epilogue, double omega,
                                        double* dot_out, void* ws, void* stream,
                                        const double* toep_host, const int* toep_rng_host,
                                        const double* dot_with, int* fused_host) {
    if (fused_host) *fused_host = 0;
    if (!x) return bad_arg(1, "x");
    if (!y) return bad_arg(2, "y");
    if (epilogue != POMS_EPI_STORE && epilogue != POMS_EPI_AXPY && !b) return bad_arg(3, "b required by epilogue");
    if (n1 < 1 || n2 < 1 || n3 < 1) return bad_arg(4, "extent");
    if (ld < n3) return bad_arg(7, "ld");
    if (pld < ld * n2) return bad_arg(8, "pld");
    if (glo < 0 || ghi < 0) return bad_arg(9, "ghost planes");
    if (!m1 || !m2 || !m3) return bad_arg(13, "band pointers");
    if (form == POMS_FORM_SUM && (!k1 || !k2 || !k3)) return bad_arg(14, "k bands");
    if (dot_out && !ws) return bad_arg(22, "ws");
    if (form != POMS_FORM_SINGLE && form != POMS_FORM_SUM) return bad_arg(12, "form");
    if (p < 1 || p > 5) return bad_arg(11, "p must be 1..5");
    MV3 a{x, y, b, n1, n2, n3, ld, pld, glo, ghi, m1, k1, m2, k2, m3, k3, omega, dot_out, ws, 0};
    if (!g_force_generic) {
        // fast path: TMA-staged pipeline (needs 16-byte aligned rows); 1 = not applicable
        const int trc = try_matvec3d_tma(a, p, form, epilogue, toep_host, toep_rng_host,
                                         (cudaStream_t)stream, dot_with, fused_host);
        if (trc == 0) {
            CHECK_LAUNCH("poms_kron_matvec_3d(tma)");
            return 0;
        }
        if (trc != 1) return trc;
    }
    const int g3 = (n3 + 63) / 64, g2 = (n2 + 15) / 16;
    a.chunk = pick_chunk(n1, (int64_t)g3 * g2, p);
    const int g1 = (n1 + a.chunk - 1) / a.chunk;
    if ((int64_t)g3 * g2 * g1 > POMS_MAX_PARTIALS) return bad_arg(4, "grid too large for ws");
    dim3 grid(g3, g2, g1);
    cudaStream_t st = (cudaStream_t)stream;
    const int rc = poms_mv3_generic_launch(a, p, form, epilogue, grid, st);
    if (rc) return rc;
    CHECK_LAUNCH("poms_kron_matvec_3d");
    return 0;
}

// ------------------------------------------------------------------------------------------
// K1: Kronecker banded mat-vec, 2-D.  CTA = 128 threads x E columns, marches along axis 1 in
// batches of RB rows per barrier.
// ------------------------------------------------------------------------------------------

template <int P, int FORM, int EPI>
__global__ void __launch_bounds__(128, 4) kron_matvec2d_kernel(MV2 a) {
    constexpr int W = 2 * P + 1;
    constexpr int NT = 128, E = 2, T2 = NT * E, RB = 8;
    constexpr int C2 = T2 + 2 * P;
    __shared__ double sx[RB][C2];
    __shared__ double red[32];
    const int i2_0 = blockIdx.x * T2;
    const int c_lo = blockIdx.y * a.chunk;
    const int c_hi = min(a.n1, c_lo + a.chunk);

    double m2c[E][W], k2c[E][W], dm[E], dk[E];
    bool v2[E];
#pragma unroll
    for (int e = 0; e < E; ++e) {
        const int i2 = i2_0 + threadIdx.x + e * NT;
        v2[e] = i2 < a.n2;
#pragma unroll
        for (int k = 0; k < W; ++k) {
            m2c[e][k] = v2[e] ? a.m2[(int64_t)i2 * W + k] : 0.0;
            k2c[e][k] = (FORM == POMS_FORM_SUM && v2[e]) ? a.k2[(int64_t)i2 * W + k] : 0.0;
        }
        dm[e] = m2c[e][P];
        dk[e] = k2c[e][P];
    }
    double acc[E][W];
#pragma unroll
    for (int e = 0; e < E; ++e)
#pragma unroll
        for (int k = 0; k < W; ++k) acc[e][k] = 0.0;
    double dsum = 0.0;

    const int start = c_lo - P, end = c_hi + P;
    int u = 0;
    for (int row0 = start; row0 < end; row0 += RB) {
        // ---- stage RB rows of x (halo'd in axis 2) ----
        for (int t = threadIdx.x; t < RB * C2; t += NT) {
            const int r = t / C2, c = t - r * C2;
            const int j1 = row0 + r, g2 = i2_0 - P + c;
            double v = 0.0;
            if (j1 < end && j1 >= -a.glo && j1 < a.n1 + a.ghi && g2 >= 0 && g2 < a.n2)
                v = __ldg(a.x + (int64_t)j1 * a.ld + g2);
            sx[r][c] = v;
        }
        __syncthreads();
        for (int r = 0; r < RB; ++r) {
            const int j1 = row0 + r;
            if (j1 >= end) break;
            double c1k[W], c1m[W], ta[E], tb[E], vout[E];
#pragma unroll
            for (int k = 0; k < W; ++k) {
                const int i1 = j1 + P - k;
                const bool ok = i1 >= 0 && i1 < a.n1;
                if (FORM == POMS_FORM_SUM) {
                    c1k[k] = ok ? __ldg(a.k1 + (int64_t)i1 * W + k) : 0.0;
                    c1m[k] = ok ? __ldg(a.m1 + (int64_t)i1 * W + k) : 0.0;
                } else {
                    c1k[k] = ok ? __ldg(a.m1 + (int64_t)i1 * W + k) : 0.0;
                    c1m[k] = 0.0;
                }
            }
#pragma unroll
            for (int e = 0; e < E; ++e) {
                const int c = threadIdx.x + e * NT;
                double uu = 0.0, vv = 0.0;
#pragma unroll
                for (int k = 0; k < W; ++k) {
                    const double xv = sx[r][c + k];
                    uu = fma(m2c[e][k], xv, uu);
                    if (FORM == POMS_FORM_SUM) vv = fma(k2c[e][k], xv, vv);
                }
                ta[e] = uu;
                tb[e] = vv;
            }
            rot_scatter<W, E, FORM == POMS_FORM_SUM>(u, acc, ta, tb, c1k, c1m, vout);
            const int i1 = j1 - P;
            if (i1 >= c_lo && i1 < c_hi) {
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    const int i2 = i2_0 + threadIdx.x + e * NT;
                    if (v2[e]) {
                        const int64_t idx = (int64_t)i1 * a.ld + i2;
                        const double v = vout[e];
                        if (EPI == POMS_EPI_STORE) {
                            a.y[idx] = v;
                            if (a.dot_out) dsum = fma(__ldg(a.x + idx), v, dsum);
                        } else if (EPI == POMS_EPI_RESID) {
                            const double rr = a.b[idx] - v;
                            a.y[idx] = rr;
                            dsum = fma(rr, rr, dsum);
                        } else if (EPI == POMS_EPI_AXPY) {
                            const double w_ = a.omega * v;
                            a.y[idx] = a.b ? a.b[idx] + w_ : w_;
                            dsum = fma(w_, w_, dsum);
                        } else {
                            double dg;
                            if (FORM == POMS_FORM_SUM)
                                dg = __ldg(a.k1 + (int64_t)i1 * W + P) * dm[e] +
                                     __ldg(a.m1 + (int64_t)i1 * W + P) * dk[e];
                            else
                                dg = __ldg(a.m1 + (int64_t)i1 * W + P) * dm[e];
                            const double dr = a.omega * (a.b[idx] - v) / dg;
                            a.y[idx] = (EPI == POMS_EPI_JACOBI) ? __ldg(a.x + idx) + dr : dr;
                            dsum = fma(dr, dr, dsum);
                        }
                    }
                }
            }
            u = (u + 1 == W) ? 0 : u + 1;
        }
        __syncthreads();
    }
    if (a.dot_out) {
        const double tot = block_sum(dsum, red);
        const unsigned nb = gridDim.x * gridDim.y;
        const unsigned bid = blockIdx.y * gridDim.x + blockIdx.x;
        grid_sum_finish(tot, a.dot_out, a.ws, nb, bid, red);
    }
}

template <int P, int FORM>
static int launch_mv2_epi(const MV2& a, int epi, dim3 grid, cudaStream_t st) {
    switch (epi) {
        case POMS_EPI_STORE: kron_matvec2d_kernel<P, FORM, POMS_EPI_STORE><<<grid, 128, 0, st>>>(a); break;
        case POMS_EPI_RESID: kron_matvec2d_kernel<P, FORM, POMS_EPI_RESID><<<grid, 128, 0, st>>>(a); break;
        case POMS_EPI_JACOBI: kron_matvec2d_kernel<P, FORM, POMS_EPI_JACOBI><<<grid, 128, 0, st>>>(a); break;
        case POMS_EPI_DINV: kron_matvec2d_kernel<P, FORM, POMS_EPI_DINV><<<grid, 128, 0, st>>>(a); break;
        case POMS_EPI_AXPY: kron_matvec2d_kernel<P, FORM, POMS_EPI_AXPY><<<grid, 128, 0, st>>>(a); break;
        default: return bad_arg(15, "epilogue");
    }
    return 0;
}
template <int P>
static int launch_mv2(const MV2& a, int form, int epi, dim3 grid, cudaStream_t st) {
    if (form == POMS_FORM_SINGLE) return launch_mv2_epi<P, POMS_FORM_SINGLE>(a, epi, grid, st);
    if (form == POMS_FORM_SUM) return launch_mv2_epi<P, POMS_FORM_SUM>(a, epi, grid, st);
    return bad_arg(10, "form");
}

extern "C" int poms_kron_matvec_2d_ex(const double* x, double* y, const double* b, int n1, int n2,
                                      int64_t ld, int glo, int ghi, int p, int form,
                                      const double* m1, const double* k1, const double* m2,
                                      const double* k2, int epilogue, double omega,
                                      double* dot_out, void* ws, void* stream,
                                      const double* toep_host, const int* toep_rng_host) {
    if (!x) return bad_arg(1, "x");
    if (!y) return bad_arg(2, "y");
    if (epilogue != POMS_EPI_STORE && epilogue != POMS_EPI_AXPY && !b) return bad_arg(3, "b required by epilogue");
    if (n1 < 1 || n2 < 1) return bad_arg(4, "extent");
    if (ld < n2) return bad_arg(6, "ld");
    if (glo < 0 || ghi < 0) return bad_arg(7, "ghost rows");
    if (!m1 || !m2) return bad_arg(11, "band pointers");
    if (form == POMS_FORM_SUM && (!k1 || !k2)) return bad_arg(12, "k bands");
    if (dot_out && !ws) return bad_arg(18, "ws");
    if (p < 1 || p > 5) return bad_arg(9, "p must be 1..5");
    if (form != POMS_FORM_SINGLE && form != POMS_FORM_SUM) return bad_arg(10, "form");
    MV2 a{x, y, b, n1, n2, ld, glo, ghi, m1, k1, m2, k2, omega, dot_out, ws, 0};
    {
        // fast path: warp-autonomous TMA kernel (16-byte aligned rows); 1 = not applicable
        const int trc = try_matvec2d_tma(a, p, form, epilogue, toep_host, toep_rng_host, (cudaStream_t)stream);
        if (trc == 0) {
            CHECK_LAUNCH("poms_kron_matvec_2d(tma)");
            return 0;
        }
        if (trc != 1) return trc;
    }
    const int g2 = (n2 + 255) / 256;
    a.chunk = pick_chunk(n1, g2, p);
    const int g1 = (n1 + a.chunk - 1) / a.chunk;
    if ((int64_t)g2 * g1 > POMS_MAX_PARTIALS) return bad_arg(4, "grid too large for ws");
    dim3 grid(g2, g1);
    cudaStream_t st = (cudaStream_t)stream;
    int rc;
    switch (p) {
        case 1: rc = launch_mv2<1>(a, form, epilogue, grid, st); break;
        case 2: rc = launch_mv2<2>(a, form, epilogue, grid, st); break;
        case 3: rc = launch_mv2<3>(a, form, epilogue, grid, st); break;
        case 4: rc = launch_mv2<4>(a, form, epilogue, grid, st); break;
        case 5: rc = launch_mv2<5>(a, form, epilogue, grid, st); break;
        default: return bad_arg(9, "p must be 1..5");
    }
    if (rc) return rc;
    CHECK_LAUNCH("poms_kron_matvec_2d");
    return 0;
}

extern "C" int poms_kron_matvec_2d(const double* x, double* y, const double* b, int n1, int n2,
                                   int64_t ld, int glo, int ghi, int p, int form,
                                   const double* m1, const double* k1, const double* m2,
                                   const double* k2, int epilogue, double omega,
                                   double* dot_out, void* ws, void* stream) {
    return poms_kron_matvec_2d_ex(x, y, b, n1, n2, ld, glo, ghi, p, form, m1, k1, m2, k2, epilogue, omega,
                                  dot_out, ws, stream, nullptr, nullptr);
}

// ------------------------------------------------------------------------------------------
// full 2-D stencil mat-vec (spl StencilMatrix.dot): coefficient-bandwidth bound
// ------------------------------------------------------------------------------------------
template <int EPI>
__global__ void __launch_bounds__(256) stencil_matvec2d_kernel(
    const double* __restrict__ x, double* __restrict__ y, const double* __restrict__ b,
    const double* __restrict__ S, int n1, int n2, int64_t ld, int glo, int ghi, int p1, int p2,
    double omega, double* dot_out, void* ws) {
    __shared__ double red[32];
    const int w1 = 2 * p1 + 1, w2 = 2 * p2 + 1;
    const int i2 = blockIdx.x * blockDim.x + threadIdx.x;
    const int i1 = blockIdx.y;
    double dsum = 0.0;
    if (i2 < n2) {
        const double* s = S + ((int64_t)i1 * n2 + i2) * (w1 * w2);
        double v = 0.0;
        for (int k1 = 0; k1 < w1; ++k1) {
            const int j1 = i1 + k1 - p1;
            if (j1 < -glo || j1 >= n1 + ghi) continue;
            for (int k2 = 0; k2 < w2; ++k2) {
                const int j2 = i2 + k2 - p2;
                if (j2 < 0 || j2 >= n2) continue;
                v = fma(s[k1 * w2 + k2], x[(int64_t)j1 * ld + j2], v);
            }
        }
        const int64_t idx = (int64_t)i1 * ld + i2;
        if (EPI == POMS_EPI_STORE) {
            y[idx] = v;
            if (dot_out) dsum = x[idx] * v;
        } else if (EPI == POMS_EPI_RESID) {
            const double rr = b[idx] - v;
            y[idx] = rr;
            dsum = rr * rr;
        } else {
            const double dr = omega * (b[idx] - v) / s[p1 * w2 + p2];
            y[idx] = (EPI == POMS_EPI_JACOBI) ? x[idx] + dr : dr;
            dsum = dr * dr;
        }
    }
    if (dot_out) {
        const double tot = block_sum(dsum, red);
        grid_sum_finish(tot, dot_out, ws, gridDim.x * gridDim.y, blockIdx.y * gridDim.x + blockIdx.x,
                        red);
    }
}

extern "C" int poms_stencil_matvec_2d(const double* x, double* y, const double* b,
                                      const double* S, int n1, int n2, int64_t ld, int glo,
                                      int ghi, int p1, int p2, int epilogue, double omega,
                                      double* dot_out, void* ws, void* stream) {
    if (!x || !y || !S) return bad_arg(1, "null pointer");
    if (epilogue != POMS_EPI_STORE && epilogue != POMS_EPI_AXPY && !b) return bad_arg(3, "b required by epilogue");
    if (dot_out && !ws) return bad_arg(15, "ws");
    dim3 grid((n2 + 255) / 256, n1);
    if ((int64_t)grid.x * grid.y > POMS_MAX_PARTIALS) return bad_arg(5, "grid too large for ws");
    cudaStream_t st = (cudaStream_t)stream;
    switch (epilogue) {
        case POMS_EPI_STORE: stencil_matvec2d_kernel<POMS_EPI_STORE><<<grid, 256, 0, st>>>(x, y, b, S, n1, n2, ld, glo, ghi, p1, p2, omega, dot_out, ws); break;
        case POMS_EPI_RESID: stencil_matvec2d_kernel<POMS_EPI_RESID><<<grid, 256, 0, st>>>(x, y, b, S, n1, n2, ld, glo, ghi, p1, p2, omega, dot_out, ws); break;
        case POMS_EPI_JACOBI: stencil_matvec2d_kernel<POMS_EPI_JACOBI><<<grid, 256, 0, st>>>(x, y, b, S, n1, n2, ld, glo, ghi, p1, p2, omega, dot_out, ws); break;
        case POMS_EPI_DINV: stencil_matvec2d_kernel<POMS_EPI_DINV><<<grid, 256, 0, st>>>(x, y, b, S, n1, n2, ld, glo, ghi, p1, p2, omega, dot_out, ws); break;
        default: return bad_arg(12, "epilogue");
    }
    CHECK_LAUNCH("poms_stencil_matvec_2d");
    return 0;
}

// ------------------------------------------------------------------------------------------
// BLAS-1 / fused CG vector kernels (grid-stride, persistent grid = 148 SMs x 8 CTAs)
// ------------------------------------------------------------------------------------------
static inline int blas_grid(int64_t n, int threads, int per_thread) {
    int64_t g = (n + (int64_t)threads * per_thread - 1) / ((int64_t)threads * per_thread);
    const int64_t cap = 148 * 8;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}

__global__ void __launch_bounds__(256) cg_update_kernel(double* __restrict__ x,
                                                        double* __restrict__ r,
                                                        const double* __restrict__ p,
                                                        const double* __restrict__ q, int64_t n,
                                                        const double* num, const double* den,
                                                        double* rr_out, void* ws) {
    __shared__ double red[32];
    const double alpha = *num / *den;
    double s = 0.0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const double pi = p[i], qi = q[i];
        x[i] = fma(alpha, pi, x[i]);
        const double ri = fma(-alpha, qi, r[i]);
        r[i] = ri;
        s = fma(ri, ri, s);
    }
    const double tot = block_sum(s, red);
    grid_sum_finish(tot, rr_out, ws, gridDim.x, blockIdx.x, red);
}

extern "C" int poms_cg_update(double* x, double* r, const double* p, const double* q, int64_t n,
                              const double* num, const double* den, double* rr_out, void* ws,
                              void* stream) {
    if (!x || !r || !p || !q) return bad_arg(1, "null vector");
    if (!num || !den || !rr_out || !ws) return bad_arg(6, "null scalar/ws");
    cg_update_kernel<<<blas_grid(n, 256, 4), 256, 0, (cudaStream_t)stream>>>(x, r, p, q, n, num,
                                                                             den, rr_out, ws);
    CHECK_LAUNCH("poms_cg_update");
    return 0;
}

__global__ void __launch_bounds__(256) p_update_kernel(double* __restrict__ p,
                                                       const double* __restrict__ s, int64_t n,
                                                       const double* num, const double* den) {
    const double beta = *num / *den;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        p[i] = fma(beta, p[i], s[i]);
}

extern "C" int poms_p_update(double* p, const double* s, int64_t n, const double* num,
                             const double* den, void* stream) {
    if (!p || !s || !num || !den) return bad_arg(1, "null pointer");
    p_update_kernel<<<blas_grid(n, 256, 4), 256, 0, (cudaStream_t)stream>>>(p, s, n, num, den);
    CHECK_LAUNCH("poms_p_update");
    return 0;
}

__global__ void __launch_bounds__(256) dot_kernel(const double* __restrict__ x,
                                                  const double* __restrict__ y, int64_t n,
                                                  double* out, void* ws) {
    __shared__ double red[32];
    double s = 0.0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        s = fma(x[i], y[i], s);
    const double tot = block_sum(s, red);
    grid_sum_finish(tot, out, ws, gridDim.x, blockIdx.x, red);
}

extern "C" int poms_dot(const double* x, const double* y, int64_t n, double* out, void* ws,
                        void* stream) {
    if (!x || !y || !out || !ws) return bad_arg(1, "null pointer");
    dot_kernel<<<blas_grid(n, 256, 4), 256, 0, (cudaStream_t)stream>>>(x, y, n, out, ws);
    CHECK_LAUNCH("poms_dot");
    return 0;
}

__global__ void __launch_bounds__(256) axpby_kernel(double* __restrict__ z, double a,
                                                    const double* __restrict__ x, double b,
                                                    const double* __restrict__ y, int64_t n) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        z[i] = y ? a * x[i] + b * y[i] : a * x[i];
}

extern "C" int poms_axpby(double* z, double a, const double* x, double b, const double* y,
                          int64_t n, void* stream) {
    if (!z || !x) return bad_arg(1, "null pointer");
    axpby_kernel<<<blas_grid(n, 256, 4), 256, 0, (cudaStream_t)stream>>>(z, a, x, b, y, n);
    CHECK_LAUNCH("poms_axpby");
    return 0;
}

__global__ void __launch_bounds__(256) axpy_dev_kernel(double* __restrict__ y,
                                                       const double* __restrict__ x, int64_t n,
                                                       const double* num, const double* den,
                                                       double sign) {
    const double a = sign * (*num / *den);
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        y[i] = fma(a, x[i], y[i]);
}

extern "C" int poms_axpy_dev(double* y, const double* x, int64_t n, const double* num,
                             const double* den, double sign, void* stream) {
    if (!y || !x || !num || !den) return bad_arg(1, "null pointer");
    axpy_dev_kernel<<<blas_grid(n, 256, 4), 256, 0, (cudaStream_t)stream>>>(y, x, n, num, den,
                                                                            sign);
    CHECK_LAUNCH("poms_axpy_dev");
    return 0;
}

__global__ void __launch_bounds__(256) cheb_update_kernel(double* __restrict__ x,
                                                          double* __restrict__ d,
                                                          const double* __restrict__ z,
                                                          double c1, double c2, int64_t n) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const double di = (c1 != 0.0) ? fma(c1, d[i], c2 * z[i]) : c2 * z[i];
        d[i] = di;
        x[i] += di;
    }
}

extern "C" int poms_cheb_update(double* x, double* d, const double* z, double c1, double c2,
                                int64_t n, void* stream) {
    if (!x || !d || !z) return bad_arg(1, "null pointer");
    cheb_update_kernel<<<blas_grid(n, 256, 4), 256, 0, (cudaStream_t)stream>>>(x, d, z, c1, c2, n);
    CHECK_LAUNCH("poms_cheb_update");
    return 0;
}

__global__ void __launch_bounds__(256) diag_scale_kernel(double* __restrict__ x,
                                                         const double* __restrict__ b,
                                                         const double* __restrict__ d, int64_t n,
                                                         double omega, double* dot_out,
                                                         void* ws) {
    __shared__ double red[32];
    double s = 0.0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const double v = omega * b[i] / d[i];
        x[i] = v;
        s = fma(v, v, s);
    }
    if (dot_out) {
        const double tot = block_sum(s, red);
        grid_sum_finish(tot, dot_out, ws, gridDim.x, blockIdx.x, red);
    }
}

extern "C" int poms_diag_scale(double* x, const double* b, const double* d, int64_t n,
                               double omega, double* dot_out, void* ws, void* stream) {
    if (!x || !b || !d) return bad_arg(1, "null pointer");
    if (dot_out && !ws) return bad_arg(7, "ws");
    diag_scale_kernel<<<blas_grid(n, 256, 4), 256, 0, (cudaStream_t)stream>>>(x, b, d, n, omega,
                                                                              dot_out, ws);
    CHECK_LAUNCH("poms_diag_scale");
    return 0;
}

// first damped-Jacobi sweep from zero / jacobi(): x = om * b / diag(A), diag from the bands
template <int FORM>
__global__ void __launch_bounds__(256) jacobi_first_kernel(
    double* __restrict__ x, const double* __restrict__ b, int n1, int n2, int n3, int64_t ld,
    int64_t pld, int W, const double* m1, const double* k1, const double* m2, const double* k2,
    const double* m3, const double* k3, double omega, double* dot_out, void* ws) {
    // 2-D is passed as n1 = 1 planes? no: 2-D uses (n1=rows -> "n2", cols -> "n3") with m1 = NULL.
    __shared__ double red[32];
    const int P = (W - 1) / 2;
    const int i3 = blockIdx.x * blockDim.x + threadIdx.x;
    const int i2 = blockIdx.y;
    double s = 0.0;
    if (i3 < n3) {
        const double m3d = m3[(int64_t)i3 * W + P], m2d = m2[(int64_t)i2 * W + P];
        double dA = m2d * m3d, dB = 0.0;
        if (FORM == POMS_FORM_SUM) dB = k2[(int64_t)i2 * W + P] * m3d + m2d * k3[(int64_t)i3 * W + P];
        for (int i1 = blockIdx.z; i1 < n1; i1 += gridDim.z) {
            double dg;
            if (m1) {
                if (FORM == POMS_FORM_SUM)
                    dg = k1[(int64_t)i1 * W + P] * dA + m1[(int64_t)i1 * W + P] * dB;
                else
                    dg = m1[(int64_t)i1 * W + P] * dA;
            } else {
                dg = (FORM == POMS_FORM_SUM) ? dB : dA;
            }
            const int64_t idx = (int64_t)i1 * pld + (int64_t)i2 * ld + i3;
            const double v = omega * b[idx] / dg;
            x[idx] = v;
            s = fma(v, v, s);
        }
    }
    if (dot_out) {
        const double tot = block_sum(s, red);
        const unsigned nb = gridDim.x * gridDim.y * gridDim.z;
        const unsigned bid = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
        grid_sum_finish(tot, dot_out, ws, nb, bid, red);
    }
}

static int jacobi_first_launch(double* x, const double* b, int n1, int n2, int n3, int64_t ld,
                               int64_t pld, int p, int form, const double* m1, const double* k1,
                               const double* m2, const double* k2, const double* m3,
                               const double* k3, double omega, double* dot_out, void* ws,
                               void* stream) {
    if (!x || !b) return bad_arg(1, "null vector");
    if (dot_out && !ws) return bad_arg(18, "ws");
    int gz = 1;
    const int gx = (n3 + 255) / 256;
    if (n1 > 1) {
        gz = (int)(POMS_MAX_PARTIALS / ((int64_t)gx * n2));
        if (gz > n1) gz = n1;
        if (gz > 64) gz = 64;
        if (gz < 1) return bad_arg(4, "grid too large for ws");
    }
    if ((int64_t)gx * n2 * gz > POMS_MAX_PARTIALS) return bad_arg(4, "grid too large for ws");
    dim3 grid(gx, n2, gz);
    cudaStream_t st = (cudaStream_t)stream;
    const int W = 2 * p + 1;
    if (form == POMS_FORM_SUM)
        jacobi_first_kernel<POMS_FORM_SUM><<<grid, 256, 0, st>>>(x, b, n1, n2, n3, ld, pld, W, m1, k1, m2, k2, m3, k3, omega, dot_out, ws);
    else
        jacobi_first_kernel<POMS_FORM_SINGLE><<<grid, 256, 0, st>>>(x, b, n1, n2, n3, ld, pld, W, m1, k1, m2, k2, m3, k3, omega, dot_out, ws);
    CHECK_LAUNCH("poms_jacobi_first");
    return 0;
}

extern "C" int poms_jacobi_first_3d(double* x, const double* b, int n1, int n2, int n3,
                                    int64_t ld, int64_t pld, int p, int form, const double* m1,
                                    const double* k1, const double* m2, const double* k2,
                                    const double* m3, const double* k3, double omega,
                                    double* dot_out, void* ws, void* stream) {
    if (!m1 || !m2 || !m3) return bad_arg(10, "bands");
    return jacobi_first_launch(x, b, n1, n2, n3, ld, pld, p, form, m1, k1, m2, k2, m3, k3, omega,
                               dot_out, ws, stream);
}

extern "C" int poms_jacobi_first_2d(double* x, const double* b, int n1, int n2, int64_t ld,
                                    int p, int form, const double* m1, const double* k1,
                                    const double* m2, const double* k2, double omega,
                                    double* dot_out, void* ws, void* stream) {
    if (!m1 || !m2) return bad_arg(8, "bands");
    // map (rows, cols) onto the (i2, i3) slots of the 3-D kernel, single plane
    return jacobi_first_launch(x, b, 1, n1, n2, ld, ld * n1, p, form, nullptr, nullptr, m1, k1,
                               m2, k2, omega, dot_out, ws, stream);
}

// ------------------------------------------------------------------------------------------
// K4: dgbtrs along one axis, one thread per line, register window
// ------------------------------------------------------------------------------------------
template <int KL, int KU>
__global__ void __launch_bounds__(128) band_solve_kernel(
    const double* __restrict__ y, double* __restrict__ x, const double* __restrict__ ab,
    const int32_t* __restrict__ ipiv, int n, int64_t n_outer, int64_t s_outer, int64_t s_axis,
    int64_t n_inner) {
    constexpr int KD = KL + KU;  // number of super-diagonals of U
    const int64_t line = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (line >= n_outer * n_inner) return;
    const int64_t o = line / n_inner, c = line - o * n_inner;
    const double* yl = y + o * s_outer + c;
    double* xl = x + o * s_outer + c;

    // ---- forward: L solve with row interchanges; window w[0..KL] = b[j..j+KL] ----
    double w[KL + 1];
#pragma unroll
    for (int m = 0; m <= KL; ++m) w[m] = m < n ? yl[(int64_t)m * s_axis] : 0.0;
    for (int j = 0; j < n; ++j) {
        if (KL > 0 && j < n - 1) {
            const int l = ipiv[j] - j;  // 0..KL
            if (l != 0) {
                const double t = w[0];
#pragma unroll
                for (int m = 1; m <= KL; ++m)
                    if (m == l) {
                        w[0] = w[m];
                        w[m] = t;
                    }
            }
            const double bj = w[0];
#pragma unroll
            for (int m = 1; m <= KL; ++m)
                if (j + m < n) w[m] = fma(-bj, __ldg(ab + (int64_t)(KD + m) * n + j), w[m]);
        }
        xl[(int64_t)j * s_axis] = w[0];
#pragma unroll
        for (int m = 0; m < KL; ++m) w[m] = w[m + 1];
        const int jn = j + KL + 1;
        w[KL] = jn < n ? yl[(int64_t)jn * s_axis] : 0.0;
    }
    // ---- backward: U x = b, U has KD super-diagonals; window u[1..KD] = x[j+1..j+KD] ----
    double u[KD + 1];
#pragma unroll
    for (int m = 0; m <= KD; ++m) u[m] = 0.0;
    for (int j = n - 1; j >= 0; --j) {
        double s = xl[(int64_t)j * s_axis];
#pragma unroll
        for (int m = KD; m >= 1; --m)
            if (j + m < n) s = fma(-__ldg(ab + (int64_t)(KD - m) * n + (j + m)), u[m], s);
        s = s / __ldg(ab + (int64_t)KD * n + j);
        xl[(int64_t)j * s_axis] = s;
#pragma unroll
        for (int m = KD; m >= 2; --m) u[m] = u[m - 1];
        if (KD >= 1) u[1] = s;
    }
}

template <int KL>
static int band_solve_ku(int ku, const double* y, double* x, const double* ab,
                         const int32_t* ipiv, int n, int64_t n_outer, int64_t s_outer,
                         int64_t s_axis, int64_t n_inner, cudaStream_t st) {
    const int64_t lines = n_outer * n_inner;
    const int grid = (int)((lines + 127) / 128);
#define BS(KU_) band_solve_kernel<KL, KU_><<<grid, 128, 0, st>>>(y, x, ab, ipiv, n, n_outer, s_outer, s_axis, n_inner)
    switch (ku) {
        case 0: BS(0); break;
        case 1: BS(1); break;
        case 2: BS(2); break;
        case 3: BS(3); break;
        case 4: BS(4); break;
        case 5: BS(5); break;
        default: return bad_arg(7, "ku must be 0..5");
    }
#undef BS
    return 0;
}


// ---- no-pivoting variants (ipiv == NULL): SPD / diagonally dominant bands (mass, GLT) ----------
// Row-oriented substitution: z_j = y_j - sum_m L[j,j-m] z_{j-m};  x_j = (z_j - sum_m U[j,j+m] x_{j+m}) / U_jj.
// The updates hit b_j in the same order as LAPACK's column-oriented dgbtrs, so the results agree
// to the last FMA.

// (a) lines along a STRIDED axis: one thread per line, lanes on the contiguous index (coalesced),
//     UNR independent loads in flight per thread.
template <int KL, int KU>
__global__ void __launch_bounds__(128) band_solve_cols_nopiv_kernel(
    const double* __restrict__ y, double* __restrict__ x, const double* __restrict__ ab, int n,
    int64_t n_outer, int64_t s_outer, int64_t s_axis, int64_t n_inner) {
    constexpr int KD = KL + KU;
    constexpr int UNR = 8;
    const int64_t line = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (line >= n_outer * n_inner) return;
    const int64_t o = line / n_inner, c = line - o * n_inner;
    const double* yl = y + o * s_outer + c;
    double* xl = x + o * s_outer + c;
    double z[KL > 0 ? KL : 1];
#pragma unroll
    for (int m = 0; m < KL; ++m) z[m] = 0.0;  // z[m-1] = z_{j-m}
    double v[UNR], vn[UNR];
#pragma unroll
    for (int u = 0; u < UNR; ++u) vn[u] = (u < n) ? yl[(int64_t)u * s_axis] : 0.0;
    for (int j0 = 0; j0 < n; j0 += UNR) {
#pragma unroll
        for (int u = 0; u < UNR; ++u) v[u] = vn[u];
        // prefetch the next batch while this one is processed
#pragma unroll
        for (int u = 0; u < UNR; ++u)
            vn[u] = (j0 + UNR + u < n) ? yl[(int64_t)(j0 + UNR + u) * s_axis] : 0.0;
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
            const int j = j0 + u;
            if (j < n) {
                double s = v[u];
#pragma unroll
                for (int m = KL; m >= 1; --m)
                    if (j - m >= 0) s = fma(-__ldg(ab + (int64_t)(KD + m) * n + (j - m)), z[m - 1], s);
                xl[(int64_t)j * s_axis] = s;
#pragma unroll
                for (int m = KL - 1; m >= 1; --m) z[m] = z[m - 1];
                if (KL > 0) z[0] = s;
            }
        }
    }
    double w[KU > 0 ? KU : 1];
#pragma unroll
    for (int m = 0; m < KU; ++m) w[m] = 0.0;  // w[m-1] = x_{j+m}
    const int nb = (n + UNR - 1) / UNR;
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
        const int j = (nb - 1) * UNR + u;
        vn[u] = (j < n) ? xl[(int64_t)j * s_axis] : 0.0;
    }
    for (int b = nb - 1; b >= 0; --b) {
        const int j0 = b * UNR;
        double rd[UNR];
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
            v[u] = vn[u];
            rd[u] = (j0 + u < n) ? 1.0 / __ldg(ab + (int64_t)KD * n + (j0 + u)) : 0.0;
        }
        if (b > 0) {
#pragma unroll
            for (int u = 0; u < UNR; ++u) vn[u] = xl[(int64_t)(j0 - UNR + u) * s_axis];
        }
#pragma unroll
        for (int u = UNR - 1; u >= 0; --u) {
            const int j = j0 + u;
            if (j < n) {
                double s = v[u];
#pragma unroll
                for (int m = KU; m >= 1; --m)
                    if (j + m < n) s = fma(-__ldg(ab + (int64_t)(KD - m) * n + (j + m)), w[m - 1], s);
                s *= rd[u];
                xl[(int64_t)j * s_axis] = s;
#pragma unroll
                for (int m = KU - 1; m >= 1; --m) w[m] = w[m - 1];
                if (KU > 0) w[0] = s;
            }
        }
    }
}

// (b) lines along the CONTIGUOUS axis: a warp owns 32 lines and walks them in 32-column tiles that
//     are staged through shared memory (cp.async double buffering), so global accesses are
//     coalesced 256-byte rows while lane l runs the recurrence of line l on the transposed tile.
#define BSR_WARPS 4
#define BSR_PITCH 33
template <int KL, int KU>
__global__ void __launch_bounds__(32 * BSR_WARPS) band_solve_rows_nopiv_kernel(
    const double* __restrict__ y, double* __restrict__ x, const double* __restrict__ ab, int n,
    int64_t n_lines, int64_t s_line, double scale, const double* add, double* out) {
    constexpr int KD = KL + KU;
    extern __shared__ double smem_bsr[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    double* tile0 = smem_bsr + (size_t)wid * 2 * 32 * BSR_PITCH;
    double* tile1 = tile0 + 32 * BSR_PITCH;
    const int64_t l0 = ((int64_t)blockIdx.x * BSR_WARPS + wid) * 32;
    if (l0 >= n_lines) return;
    const int nl = (int)min((int64_t)32, n_lines - l0);
    const int nt = (n + 31) / 32;
    const bool mine = lane < nl;

    auto prefetch = [&](const double* src, double* tile, int t) {
        const int col = t * 32 + lane;
        if (col < n) {
            for (int r = 0; r < nl; ++r) cp_async8(tile + r * BSR_PITCH + lane, src + (l0 + r) * s_line + col);
        }
        cp_async_commit();
    };
    auto store = [&](double* tile, int t) {
        const int col = t * 32 + lane;
        if (col < n) {
            for (int r = 0; r < nl; ++r) x[(l0 + r) * s_line + col] = tile[r * BSR_PITCH + lane];
        }
    };
    // final store of the backward sweep: out = [add +] scale * solution (fused smoother update)
    auto store_final = [&](double* tile, int t) {
        const int col = t * 32 + lane;
        if (col < n) {
            if (add) {
                // batches of 8 rows: all loads first (add may alias out, so the compiler cannot
                // hoist them across the stores itself)
                for (int r0 = 0; r0 < nl; r0 += 8) {
                    double a8[8];
#pragma unroll
                    for (int q = 0; q < 8; ++q)
                        a8[q] = (r0 + q < nl) ? add[(l0 + r0 + q) * s_line + col] : 0.0;
#pragma unroll
                    for (int q = 0; q < 8; ++q)
                        if (r0 + q < nl)
                            out[(l0 + r0 + q) * s_line + col] =
                                __dadd_rn(a8[q], __dmul_rn(scale, tile[(r0 + q) * BSR_PITCH + lane]));
                }
            } else {
                for (int r = 0; r < nl; ++r)
                    out[(l0 + r) * s_line + col] = __dmul_rn(scale, tile[r * BSR_PITCH + lane]);
            }
        }
    };

    // ---------------- forward ----------------
    double z[KL > 0 ? KL : 1];
#pragma unroll
    for (int m = 0; m < KL; ++m) z[m] = 0.0;
    prefetch(y, tile0, 0);
    for (int t = 0; t < nt; ++t) {
        double* cur = (t & 1) ? tile1 : tile0;
        double* nxt = (t & 1) ? tile0 : tile1;
        if (t + 1 < nt) prefetch(y, nxt, t + 1); else cp_async_commit();
        cp_async_wait<1>();
        __syncwarp();
        const int jend = min(32, n - t * 32);
        if (mine) {
            double* row = cur + lane * BSR_PITCH;
            for (int u0 = 0; u0 < jend; u0 += 8) {
                double c[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) c[q] = row[min(u0 + q, 31)];
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const int j = t * 32 + u0 + q;
                    if (u0 + q < jend) {
                        double s = c[q];
#pragma unroll
                        for (int m = KL; m >= 1; --m)
                            if (j - m >= 0)
                                s = fma(-__ldg(ab + (int64_t)(KD + m) * n + (j - m)), z[m - 1], s);
                        c[q] = s;
#pragma unroll
                        for (int m = KL - 1; m >= 1; --m) z[m] = z[m - 1];
                        if (KL > 0) z[0] = s;
                    }
                }
#pragma unroll
                for (int q = 0; q < 8; ++q)
                    if (u0 + q < jend) row[u0 + q] = c[q];
            }
        }
        __syncwarp();
        store(cur, t);
        __syncwarp();
    }
    cp_async_wait<0>();
    __syncwarp();
    // ---------------- backward (reads the forward result back from x) ----------------
    double w[KU > 0 ? KU : 1];
#pragma unroll
    for (int m = 0; m < KU; ++m) w[m] = 0.0;
    prefetch(x, tile0, nt - 1);
    for (int i = 0; i < nt; ++i) {
        const int t = nt - 1 - i;
        double* cur = (i & 1) ? tile1 : tile0;
        double* nxt = (i & 1) ? tile0 : tile1;
        if (t - 1 >= 0) prefetch(x, nxt, t - 1); else cp_async_commit();
        cp_async_wait<1>();
        __syncwarp();
        const int jend = min(32, n - t * 32);
        if (mine) {
            double* row = cur + lane * BSR_PITCH;
            for (int u0 = 24; u0 >= 0; u0 -= 8) {
                if (u0 >= jend) continue;
                double c[8], rd[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    c[q] = row[u0 + q];
                    rd[q] = 1.0 / __ldg(ab + (int64_t)KD * n + min(t * 32 + u0 + q, n - 1));
                }
#pragma unroll
                for (int q = 7; q >= 0; --q) {
                    const int j = t * 32 + u0 + q;
                    if (u0 + q < jend) {
                        double s = c[q];
#pragma unroll
                        for (int m = KU; m >= 1; --m)
                            if (j + m < n)
                                s = fma(-__ldg(ab + (int64_t)(KD - m) * n + (j + m)), w[m - 1], s);
                        s *= rd[q];
                        c[q] = s;
#pragma unroll
                        for (int m = KU - 1; m >= 1; --m) w[m] = w[m - 1];
                        if (KU > 0) w[0] = s;
                    }
                }
#pragma unroll
                for (int q = 0; q < 8; ++q)
                    if (u0 + q < jend) row[u0 + q] = c[q];
            }
        }
        __syncwarp();
        store_final(cur, t);
        __syncwarp();
    }
    cp_async_wait<0>();
}

template <int KL>
static int band_solve_nopiv_ku(int ku, const double* y, double* x, const double* ab, int n,
                               int64_t n_outer, int64_t s_outer, int64_t s_axis, int64_t n_inner,
                               cudaStream_t st, double scale = 1.0, const double* add = nullptr,
                               double* out = nullptr) {
    const bool rows = (n_inner == 1 && s_axis == 1);
    if (!out) out = x;
    if (!rows && (add || out != x || scale != 1.0)) return bad_arg(13, "fused epilogue needs the contiguous axis");
    const int64_t lines = n_outer * n_inner;
    const size_t smem = (size_t)BSR_WARPS * 2 * 32 * BSR_PITCH * sizeof(double);
#define BSN(KU_)                                                                                  \
    if (rows) {                                                                                   \
        static bool attr_set = false;                                                             \
        if (!attr_set) {                                                                          \
            cudaFuncSetAttribute(band_solve_rows_nopiv_kernel<KL, KU_>,                           \
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);         \
            attr_set = true;                                                                      \
        }                                                                                         \
        const int grid = (int)((lines + 32 * BSR_WARPS - 1) / (32 * BSR_WARPS));                  \
        band_solve_rows_nopiv_kernel<KL, KU_><<<grid, 32 * BSR_WARPS, smem, st>>>(                \
            y, x, ab, n, lines, s_outer, scale, add, out);                                        \
    } else {                                                                                      \
        const int grid = (int)((lines + 127) / 128);                                              \
        band_solve_cols_nopiv_kernel<KL, KU_><<<grid, 128, 0, st>>>(y, x, ab, n, n_outer,         \
                                                                     s_outer, s_axis, n_inner);    \
    }
    switch (ku) {
        case 0: BSN(0); break;
        case 1: BSN(1); break;
        case 2: BSN(2); break;
        case 3: BSN(3); break;
        case 4: BSN(4); break;
        case 5: BSN(5); break;
        default: return bad_arg(7, "ku must be 0..5");
    }
#undef BSN
    return 0;
}

// (c) few lines (2-D grids): CHUNKED substitution.  For the diagonally dominant SPD bands of this
//     code the homogeneous solutions of both triangular recurrences decay geometrically, so a line
//     can be cut into chunks that each start `warm` entries early from a zero state: after the
//     warm-up the state agrees with the sequential sweep to rounding (the caller verifies the decay
//     for the given factor before choosing this path).  One thread per (line, chunk); the forward
//     and backward sweeps are separate launches (y -> work, work -> x).
template <int KL, int KU, bool FWD>
__global__ void __launch_bounds__(128) band_chunk_kernel(
    const double* __restrict__ in, double* __restrict__ out, const double* __restrict__ ab, int n,
    int64_t n_outer, int64_t s_outer, int64_t s_axis, int64_t n_inner, int chunk, int warm,
    int nchunks) {
    constexpr int KD = KL + KU;
    const int64_t nlines = n_outer * n_inner;
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nlines * nchunks) return;
    const int ck = (int)(t / nlines);
    const int64_t line = t - (int64_t)ck * nlines;
    const int64_t o = line / n_inner, c = line - o * n_inner;
    const double* il = in + o * s_outer + c;
    double* ol = out + o * s_outer + c;
    const int jb = ck * chunk, je = min(n, jb + chunk);
    if (FWD) {
        double z[KL > 0 ? KL : 1];
#pragma unroll
        for (int m = 0; m < KL; ++m) z[m] = 0.0;
        const int j0 = max(0, jb - warm);
        for (int j = j0; j < je; ++j) {
            double s = il[(int64_t)j * s_axis];
#pragma unroll
            for (int m = KL; m >= 1; --m)
                if (j - m >= j0) s = fma(-__ldg(ab + (int64_t)(KD + m) * n + (j - m)), z[m - 1], s);
            if (j >= jb) ol[(int64_t)j * s_axis] = s;
#pragma unroll
            for (int m = KL - 1; m >= 1; --m) z[m] = z[m - 1];
            if (KL > 0) z[0] = s;
        }
    } else {
        double w[KU > 0 ? KU : 1];
#pragma unroll
        for (int m = 0; m < KU; ++m) w[m] = 0.0;
        const int jt = min(n, je + warm);
        for (int j = jt - 1; j >= jb; --j) {
            double s = il[(int64_t)j * s_axis];
#pragma unroll
            for (int m = KU; m >= 1; --m)
                if (j + m < jt) s = fma(-__ldg(ab + (int64_t)(KD - m) * n + (j + m)), w[m - 1], s);
            s *= 1.0 / __ldg(ab + (int64_t)KD * n + j);
            if (j < je) ol[(int64_t)j * s_axis] = s;
#pragma unroll
            for (int m = KU - 1; m >= 1; --m) w[m] = w[m - 1];
            if (KU > 0) w[0] = s;
        }
    }
}

template <int KL>
static int band_chunk_ku(int ku, const double* y, double* x, double* work, const double* ab, int n,
                         int64_t n_outer, int64_t s_outer, int64_t s_axis, int64_t n_inner,
                         int chunk, int warm_f, int warm_b, cudaStream_t st) {
    const int nchunks = (n + chunk - 1) / chunk;
    const int64_t threads = n_outer * n_inner * nchunks;
    const int grid = (int)((threads + 127) / 128);
#define BC(KU_)                                                                                   \
    band_chunk_kernel<KL, KU_, true><<<grid, 128, 0, st>>>(y, work, ab, n, n_outer, s_outer,      \
                                                           s_axis, n_inner, chunk, warm_f, nchunks); \
    band_chunk_kernel<KL, KU_, false><<<grid, 128, 0, st>>>(work, x, ab, n, n_outer, s_outer,     \
                                                            s_axis, n_inner, chunk, warm_b, nchunks);
    switch (ku) {
        case 0: BC(0); break;
        case 1: BC(1); break;
        case 2: BC(2); break;
        case 3: BC(3); break;
        case 4: BC(4); break;
        case 5: BC(5); break;
        default: return bad_arg(7, "ku must be 0..5");
    }
#undef BC
    return 0;
}

extern "C" int poms_band_solve_axis_chunked(const double* y, double* x, double* work,
                                            const double* ab, int n, int kl, int ku,
                                            int64_t n_outer, int64_t s_outer, int64_t s_axis,
                                            int64_t n_inner, int chunk, int warm_fwd, int warm_bwd,
                                            void* stream) {
    if (!y || !x || !work || !ab) return bad_arg(1, "null pointer");
    if (work == y || work == x) return bad_arg(3, "work must not alias y or x");
    if (n < 1 || n_outer < 1 || n_inner < 1) return bad_arg(5, "extent");
    if (chunk < 8 || warm_fwd < 0 || warm_bwd < 0) return bad_arg(12, "chunk/warm");
    cudaStream_t st = (cudaStream_t)stream;
    int rc;
    switch (kl) {
        case 0: rc = band_chunk_ku<0>(ku, y, x, work, ab, n, n_outer, s_outer, s_axis, n_inner, chunk, warm_fwd, warm_bwd, st); break;
        case 1: rc = band_chunk_ku<1>(ku, y, x, work, ab, n, n_outer, s_outer, s_axis, n_inner, chunk, warm_fwd, warm_bwd, st); break;
        case 2: rc = band_chunk_ku<2>(ku, y, x, work, ab, n, n_outer, s_outer, s_axis, n_inner, chunk, warm_fwd, warm_bwd, st); break;
        case 3: rc = band_chunk_ku<3>(ku, y, x, work, ab, n, n_outer, s_outer, s_axis, n_inner, chunk, warm_fwd, warm_bwd, st); break;
        case 4: rc = band_chunk_ku<4>(ku, y, x, work, ab, n, n_outer, s_outer, s_axis, n_inner, chunk, warm_fwd, warm_bwd, st); break;
        case 5: rc = band_chunk_ku<5>(ku, y, x, work, ab, n, n_outer, s_outer, s_axis, n_inner, chunk, warm_fwd, warm_bwd, st); break;
        default: return bad_arg(6, "kl must be 0..5");
    }
    if (rc) return rc;
    g_launches++;   // two launches
    CHECK_LAUNCH("poms_band_solve_axis_chunked");
    return 0;
}

extern "C" int poms_band_solve_axis_fused(const double* y, double* work, const double* ab, int n,
                                          int kl, int ku, int64_t n_lines, int64_t s_line,
                                          double scale, const double* add, double* out,
                                          void* stream) {
    if (!y || !work || !ab || !out) return bad_arg(1, "null pointer");
    if (n < 1 || n_lines < 1) return bad_arg(4, "extent");
    cudaStream_t st = (cudaStream_t)stream;
    int rc;
    switch (kl) {
        case 0: rc = band_solve_nopiv_ku<0>(ku, y, work, ab, n, n_lines, s_line, 1, 1, st, scale, add, out); break;
        case 1: rc = band_solve_nopiv_ku<1>(ku, y, work, ab, n, n_lines, s_line, 1, 1, st, scale, add, out); break;
        case 2: rc = band_solve_nopiv_ku<2>(ku, y, work, ab, n, n_lines, s_line, 1, 1, st, scale, add, out); break;
        case 3: rc = band_solve_nopiv_ku<3>(ku, y, work, ab, n, n_lines, s_line, 1, 1, st, scale, add, out); break;
        case 4: rc = band_solve_nopiv_ku<4>(ku, y, work, ab, n, n_lines, s_line, 1, 1, st, scale, add, out); break;
        case 5: rc = band_solve_nopiv_ku<5>(ku, y, work, ab, n, n_lines, s_line, 1, 1, st, scale, add, out); break;
        default: return bad_arg(5, "kl must be 0..5");
    }
    if (rc) return rc;
    CHECK_LAUNCH("poms_band_solve_axis_fused");
    return 0;
}

extern "C" int poms_band_solve_axis(const double* y, double* x, const double* ab,
                                    const int32_t* ipiv, int n, int kl, int ku, int64_t n_outer,
                                    int64_t s_outer, int64_t s_axis, int64_t n_inner,
                                    void* stream) {
    if (!y || !x || !ab) return bad_arg(1, "null pointer");
    if (n < 1 || n_outer < 1 || n_inner < 1) return bad_arg(5, "extent");
    cudaStream_t st = (cudaStream_t)stream;
    int rc;
    if (!ipiv) {  // factorisation without row interchanges
        switch (kl) {
            case 0: rc = band_solve_nopiv_ku<0>(ku, y, x, ab, n, n_outer, s_outer, s_axis, n_inner, st); break;
            case 1: rc = band_solve_nopiv_ku<1>(ku, y, x, ab, n, n_outer, s_outer, s_axis, n_inner, st); break;
            case 2: rc = band_solve_nopiv_ku<2>(ku, y, x, ab, n, n_outer, s_outer, s_axis, n_inner, st); break;
            case 3: rc = band_solve_nopiv_ku<3>(ku, y, x, ab, n, n_outer, s_outer, s_axis, n_inner, st); break;
            case 4: rc = band_solve_nopiv_ku<4>(ku, y, x, ab, n, n_outer, s_outer, s_axis, n_inner, st); break;
            case 5: rc = band_solve_nopiv_ku<5>(ku, y, x, ab, n, n_outer, s_outer, s_axis, n_inner, st); break;
            default: return bad_arg(6, "kl must be 0..5");
        }
        if (rc) return rc;
        CHECK_LAUNCH("poms_band_solve_axis(nopiv)");
        return 0;
    }
    switch (kl) {
        case 0: rc = band_solve_ku<0>(ku, y, x, ab, ipiv, n, n_outer, s_outer, s_axis, n_inner, st); break;
        case 1: rc = band_solve_ku<1>(ku, y, x, ab, ipiv, n, n_outer, s_outer, s_axis, n_inner, st); break;
        case 2: rc = band_solve_ku<2>(ku, y, x, ab, ipiv, n, n_outer, s_outer, s_axis, n_inner, st); break;
        case 3: rc = band_solve_ku<3>(ku, y, x, ab, ipiv, n, n_outer, s_outer, s_axis, n_inner, st); break;
        case 4: rc = band_solve_ku<4>(ku, y, x, ab, ipiv, n, n_outer, s_outer, s_axis, n_inner, st); break;
        case 5: rc = band_solve_ku<5>(ku, y, x, ab, ipiv, n, n_outer, s_outer, s_axis, n_inner, st); break;
        default: return bad_arg(6, "kl must be 0..5");
    }
    if (rc) return rc;
    CHECK_LAUNCH("poms_band_solve_axis");
    return 0;
}

// ------------------------------------------------------------------------------------------
// K5: per-axis sparse row gather (prolongation / restriction along one axis)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) axis_gather_kernel(
    const double* __restrict__ in, double* __restrict__ out, const int32_t* __restrict__ start,
    const double* __restrict__ coef, int W, int n_in, int n_out, int64_t n_outer, int64_t so_in,
    int64_t sa_in, int64_t so_out, int64_t sa_out, int64_t n_inner, int accumulate) {
    const int64_t total = n_outer * n_out * n_inner;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
        const int64_t c = t % n_inner;
        const int64_t oi = t / n_inner;
        const int i = (int)(oi % n_out);
        const int64_t o = oi / n_out;
        const int s0 = start[i];
        const double* ip = in + o * so_in + c;
        double v = 0.0;
        for (int w = 0; w < W; ++w) {
            const int j = s0 + w;
            if (j >= 0 && j < n_in) v = fma(__ldg(coef + (int64_t)i * W + w), ip[(int64_t)j * sa_in], v);
        }
        double* op = out + o * so_out + (int64_t)i * sa_out + c;
        *op = accumulate ? *op + v : v;
    }
}

// Fast path for passes along a NON-contiguous axis: one output row per blockIdx.y, lanes on the
// contiguous index with 128-bit loads/stores, the W coefficients of the row are block-uniform.
template <int ACC>
__global__ void __launch_bounds__(256) axis_gather_strided2_kernel(
    const double2* __restrict__ in, double2* __restrict__ out, const int32_t* __restrict__ start,
    const double* __restrict__ coef, int W, int n_in, int n_out, int64_t so_in2, int64_t sa_in2,
    int64_t so_out2, int64_t sa_out2, int64_t n_inner2) {
    // One output row (plane) per blockIdx.y: its <= 8 taps are block-uniform.  Taps with a ZERO
    // coefficient are skipped (knot-insertion rows carry p+1 slots of which 2-3 are used: a quarter to
    // a half of the input planes used to be read for nothing), and every thread works on two column
    // pairs at once (two independent load streams per tap).
    const int i = blockIdx.y;
    const int64_t o = blockIdx.z;
    const int s0 = __ldg(start + i);
    const double* cf = coef + (int64_t)i * W;
    const double2* ip = in + o * so_in2;
    double2* op = out + o * so_out2 + (int64_t)i * sa_out2;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < n_inner2; c += 2 * stride) {
        const int64_t c1 = c + stride;
        const bool ok1 = c1 < n_inner2;
        double2 old0 = make_double2(0.0, 0.0), old1 = make_double2(0.0, 0.0);
        if (ACC) {
            old0 = op[c];
            if (ok1) old1 = op[c1];
        }
        double2 v0 = make_double2(0.0, 0.0), v1 = make_double2(0.0, 0.0);
#pragma unroll
        for (int w = 0; w < 8; ++w) {
            if (w < W) {
                const int j = s0 + w;
                const double cw = __ldg(cf + w);
                if (cw != 0.0 && j >= 0 && j < n_in) {
                    const double2* row = ip + (int64_t)j * sa_in2;
                    const double2 x0 = row[c];
                    const double2 x1 = ok1 ? row[c1] : make_double2(0.0, 0.0);
                    v0.x = fma(cw, x0.x, v0.x);
                    v0.y = fma(cw, x0.y, v0.y);
                    v1.x = fma(cw, x1.x, v1.x);
                    v1.y = fma(cw, x1.y, v1.y);
                }
            }
        }
        if (ACC) {
            v0.x += old0.x;
            v0.y += old0.y;
            v1.x += old1.x;
            v1.y += old1.y;
        }
        op[c] = v0;
        if (ok1) op[c1] = v1;
    }
}

extern "C" int poms_axis_gather(const double* in, double* out, const int32_t* start,
                                const double* coef, int W, int n_in, int n_out, int64_t n_outer,
                                int64_t so_in, int64_t sa_in, int64_t so_out, int64_t sa_out,
                                int64_t n_inner, int accumulate, void* stream) {
    if (!in || !out || !start || !coef) return bad_arg(1, "null pointer");
    if (W < 1 || n_in < 1 || n_out < 1 || n_outer < 1 || n_inner < 1) return bad_arg(5, "extent");
    const bool even = !((n_inner | so_in | sa_in | so_out | sa_out) & 1) &&
                      !(((uintptr_t)in | (uintptr_t)out) & 15);
    if (even && n_inner >= 64 && n_out <= 65535 && n_outer <= 65535 && W <= 8) {
        const int64_t n2 = n_inner / 2;
        int gx = (int)((n2 + 511) / 512);   // ~2 double2 per thread
        if (gx < 1) gx = 1;
        dim3 grid(gx, n_out, (unsigned)n_outer);
        if (accumulate)
            axis_gather_strided2_kernel<1><<<grid, 256, 0, (cudaStream_t)stream>>>(
                (const double2*)in, (double2*)out, start, coef, W, n_in, n_out, so_in / 2, sa_in / 2,
                so_out / 2, sa_out / 2, n2);
        else
            axis_gather_strided2_kernel<0><<<grid, 256, 0, (cudaStream_t)stream>>>(
                (const double2*)in, (double2*)out, start, coef, W, n_in, n_out, so_in / 2, sa_in / 2,
                so_out / 2, sa_out / 2, n2);
        CHECK_LAUNCH("poms_axis_gather(strided2)");
        return 0;
    }
    const int64_t total = n_outer * n_out * n_inner;
    int64_t g = (total + 255) / 256;
    if (g > 148 * 16) g = 148 * 16;
    axis_gather_kernel<<<(int)g, 256, 0, (cudaStream_t)stream>>>(
        in, out, start, coef, W, n_in, n_out, n_outer, so_in, sa_in, so_out, sa_out, n_inner,
        accumulate);
    CHECK_LAUNCH("poms_axis_gather");
    return 0;
}

// ------------------------------------------------------------------------------------------
// dense mat-vec for the replicated coarse solve: one warp per row
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) dense_matvec_kernel(const double* __restrict__ A,
                                                           const double* __restrict__ x,
                                                           double* __restrict__ y, int n) {
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= n) return;
    const double* a = A + (int64_t)row * n;
    double s = 0.0;
    for (int j = lane; j < n; j += 32) s = fma(a[j], x[j], s);
    s = warp_sum(s);
    if (lane == 0) y[row] = s;
}

extern "C" int poms_dense_matvec(const double* Ainv, const double* x, double* y, int n,
                                 void* stream) {
    if (!Ainv || !x || !y) return bad_arg(1, "null pointer");
    if (n < 1) return bad_arg(4, "n");
    const int grid = (n * 32 + 255) / 256;
    dense_matvec_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(Ainv, x, y, n);
    CHECK_LAUNCH("poms_dense_matvec");
    return 0;
}

#include "poms_transfer3d.cuh"
#endif  // POMS_TU == 0
