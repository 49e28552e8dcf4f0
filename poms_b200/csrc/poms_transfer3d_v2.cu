// poms_transfer3d_v2.cu -- one-pass 3-D knot-insertion transfers for the BIG levels (round 2, late).
// Self-contained unit of libpoms_b200.so (own kernels + C ABI, see include/poms_b200.h).
//
//   poms_prolong_3d_v2:  fine (+)= (P1 (x) P2 (x) P3) coarse   -- P.dot + correction,
//                                                /root/reference/sources/mg_jac.py:102,112
//   poms_restrict_3d_v2: coarse = (R1 (x) R2 (x) R3) fine       -- R.dot, mg_jac.py:94
//
// Same rows (poms_axis_gather format: out[i] = sum_{t<W} coef[i*W+t] * in[start[i]+t], nondecreasing
// starts, zero coefficients on taps outside the input) and same tiles as the round-1 one-pass kernels
// in poms_transfer3d.cuh, which lose to three per-axis passes at 515^3 (0.77 / 1.41 ms against
// 0.62 / 0.96 ms) although they move 9 / 17 instead of 21 / 29 bytes per fine point.  What changes:
//   * the in-plane (axes 2, 3) passes through shared memory run once per COARSE plane, not once per
//     fine plane: prolongation applies them to the coarse plane BEFORE the axis-1 combination,
//     restriction AFTER it -- half the barriers and half the shared-memory work per fine point;
//   * the axis-1 pass lives in registers: a sliding window of W in-plane-prolonged coarse planes
//     (gather form) resp. the partial sums of the <= NS coarse planes open at a time (scatter form);
//   * fine-grid traffic goes global <-> registers directly, two planes ahead of its use (no cp.async
//     staging, no shared-memory round trip of the fine tile in the prolongation);
//   * the row width W is a template parameter and the tap loops carry no bounds checks: the tile
//     buffers are zero-initialised with a pad row, so a tap that leaves the tile meets a finite value
//     and, by the row format, a zero coefficient.
//
// POMS_HOST_EMU is never defined by build.py: libpoms_b200.so contains the CUDA kernels only.  The
// CPU test tests/test_kernel_host_emulation.py defines it to compile THIS file for the host over
// tests/host_emu/cuda_emu.h (one OS thread per CUDA thread, a pthread barrier for __syncthreads) under
// AddressSanitizer and ThreadSanitizer: out-of-bounds accesses and missing barriers of the kernels
// below show up without a GPU, on every row width and on ragged / slab-plan shapes.
#ifdef POMS_HOST_EMU
#include "cuda_emu.h"
#else
#define POMS_TU 99
#include "poms_kernels.cu"
#endif

namespace {

struct TR3 {
    const double* src;   // prolongation: coarse;  restriction: fine
    double* dst;         // prolongation: fine;    restriction: coarse
    int n1f, n2f, n3f, n1c, n2c, n3c;
    int64_t ldf, pldf, ldc, pldc;
    const int32_t *s1, *s2, *s3;
    const double *c1, *c2, *c3;
    int chunk, accumulate;
};

int t_bad_arg(int idx, const char* what) {
    snprintf(g_err, sizeof(g_err), "bad argument %d: %s", idx, what);
    return -idx;
}

// largest input extent of a tile of `tile` consecutive output rows; -1 if the starts decrease
int rows_extent(const int32_t* s, int n_out, int W, int n_in, int tile, bool any_start = false) {
    int ext = 0;
    for (int i = 1; i < n_out; ++i)
        if (s[i] < s[i - 1]) return -1;
    for (int i0 = 0; i0 < n_out; i0 += tile) {
        const int i1 = (i0 + tile < n_out ? i0 + tile : n_out) - 1;
        int hi = s[i1] + W - 1;
        if (hi > n_in - 1) hi = n_in - 1;
        if (!any_start && (s[i0] < 0 || s[i0] > n_in - 1)) return -1;
        if (hi - s[i0] + 1 > ext) ext = hi - s[i0] + 1;
    }
    return ext;
}
// largest number of rows whose tap range [s, s+W-1] contains the same input index
int rows_open(const int32_t* s, int n_out, int W) {
    int best = 0, lo = 0;
    for (int i = 0; i < n_out; ++i) {
        while (s[lo] + W - 1 < s[i]) ++lo;
        if (i - lo + 1 > best) best = i - lo + 1;
    }
    return best;
}

// ---------------------------------------------------------------------------------------------
// Prolongation.  CTA: 16 x 64 fine tile of axes (2, 3), marching along a chunk of FINE planes.
//   per coarse plane c (once):  G = coarse tile (global -> shared, prefetched one plane ahead)
//                               H = G prolonged along axis 3          (shared -> shared)
//                               g[W-1] = H prolonged along axis 2     (shared -> registers)
//   per fine plane j:           x[j] (+)= sum_k P1[j][k] * g[k]       (registers; x two planes ahead)
// ---------------------------------------------------------------------------------------------
constexpr int P2_F2 = 16, P2_F3 = 64, P2_RC2 = 16, P2_RC3 = 40, P2_MAXCH = 64;
constexpr int P2_GS = P2_RC3 + 1;
// coarse-tile points per thread: the true extent nr2 x nr3 is <= 256 * p2_ng(W) (host-checked)
__host__ __device__ constexpr int p2_ng(int W) { return W >= 6 ? 3 : 2; }

template <int W>
__global__ void __launch_bounds__(256, 3) prolong3d_v2_kernel(const TR3 a) {
    __shared__ double G[(P2_RC2 + 1) * P2_GS];
    __shared__ double H[(P2_RC2 + 8) * P2_F3];
    __shared__ double c2s[P2_F2 * W];
    __shared__ double c1s[P2_MAXCH * W];
    __shared__ double c3s[P2_F3 * (W | 1)];    // odd row stride: conflict-free per-lane rows
    __shared__ int s1s[P2_MAXCH];
    constexpr int WS = W | 1, P2_NG = p2_ng(W);
    const int tid = threadIdx.x;
    const int tx = tid & (P2_F3 - 1), ty = tid >> 6;
    const int f3_0 = blockIdx.x * P2_F3, f2_0 = blockIdx.y * P2_F2;
    const int j_lo = blockIdx.z * a.chunk, j_hi = min(a.n1f, j_lo + a.chunk);
    const int f3l = min(P2_F3, a.n3f - f3_0), f2l = min(P2_F2, a.n2f - f2_0);
    const int c2lo = a.s2[f2_0], c3lo = a.s3[f3_0];
    const int c2hi = min(a.n2c - 1, a.s2[f2_0 + f2l - 1] + W - 1);
    const int c3hi = min(a.n3c - 1, a.s3[f3_0 + f3l - 1] + W - 1);
    const int nr2 = c2hi - c2lo + 1, nr3 = c3hi - c3lo + 1;

    for (int t = tid; t < (P2_RC2 + 1) * P2_GS; t += 256) G[t] = 0.0;
    for (int t = tid; t < (P2_RC2 + 8) * P2_F3; t += 256) H[t] = 0.0;
    for (int t = tid; t < P2_F2 * W; t += 256) {
        const int r = t / W;
        c2s[t] = r < f2l ? a.c2[(int64_t)f2_0 * W + t] : 0.0;
    }
    for (int t = tid; t < (j_hi - j_lo) * W; t += 256) c1s[t] = a.c1[(int64_t)j_lo * W + t];
    for (int t = tid; t < j_hi - j_lo; t += 256) s1s[t] = a.s1[j_lo + t];

    const bool v3 = tx < f3l;
    for (int t = tid; t < P2_F3 * W; t += 256) {
        const int c = t / W, k = t - c * W;
        c3s[c * WS + k] = c < f3l ? a.c3[(int64_t)f3_0 * W + t] : 0.0;
    }
    const int o3 = v3 ? a.s3[f3_0 + tx] - c3lo : 0;
    int o2r[4];
    unsigned okmask = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int r = ty + 4 * q;
        o2r[q] = r < f2l ? a.s2[f2_0 + r] - c2lo : 0;
        if (v3 && r < f2l) okmask |= 1u << q;
    }
    // this thread's slots of the coarse tile
    int goff[P2_NG], gidx[P2_NG];
    unsigned gmask = 0;
#pragma unroll
    for (int m = 0; m < P2_NG; ++m) {
        const int slot = tid + 256 * m;
        const int r = slot / nr3, c = slot - r * nr3;
        const bool ok = r < nr2;
        if (ok) gmask |= 1u << m;
        goff[m] = ok ? (int)((int64_t)(c2lo + r) * a.ldc) + c3lo + c : 0;
        gidx[m] = ok ? r * P2_GS + c : 0;
    }
    __syncthreads();

    int cnext = s1s[0];          // coarse plane that enters the window next
    int cb = cnext - W;          // coarse plane held in g[0]
    double gpre[P2_NG];
    auto load_coarse = [&](const int c) {
        const double* const pl = a.src + (int64_t)max(0, min(c, a.n1c - 1)) * a.pldc;
#pragma unroll
        for (int m = 0; m < P2_NG; ++m) gpre[m] = (gmask >> m & 1u) ? __ldg(pl + goff[m]) : 0.0;
    };
    load_coarse(cnext);

    double* xp = a.dst + (int64_t)j_lo * a.pldf + (int64_t)(f2_0 + ty) * a.ldf + f3_0 + tx;
    const int64_t rstep = 4 * a.ldf;
    // x of the planes j and j+1 (a third plane in flight at two CTAs per SM, without the 20 bytes of
    // spills of this version, measured slower: 0.80 vs 0.72 ms at 515^3)
    double xa[4], xb[4];
    auto load_x = [&](double (&x)[4], const double* p, const bool live) {
#pragma unroll
        for (int q = 0; q < 4; ++q)
            x[q] = (live && (okmask >> q & 1u)) ? __ldcs(p + q * rstep) : 0.0;
    };
    const bool acc_in = a.accumulate != 0;
    load_x(xa, xp, acc_in);
    load_x(xb, xp + a.pldf, acc_in && j_lo + 1 < j_hi);

    double g[W][4];
#pragma unroll
    for (int k = 0; k < W; ++k)
#pragma unroll
        for (int q = 0; q < 4; ++q) g[k][q] = 0.0;

#pragma unroll 1
    for (int j = j_lo; j < j_hi; ++j) {
        const int p0 = s1s[j - j_lo];
#pragma unroll 1
        while (cb < p0) {
#pragma unroll
            for (int k = 0; k + 1 < W; ++k)
#pragma unroll
                for (int q = 0; q < 4; ++q) g[k][q] = g[k + 1][q];
#pragma unroll
            for (int m = 0; m < P2_NG; ++m)
                if (gmask >> m & 1u) G[gidx[m]] = gpre[m];
            ++cnext;
            load_coarse(cnext);
            __syncthreads();
            if (v3) {
#pragma unroll
                for (int i = 0; i < P2_RC2 / 4; ++i) {
                    const int r = ty + 4 * i;
                    if (r < nr2) {
                        const double* const gr = G + r * P2_GS + o3;
                        const double* const c3r = c3s + tx * WS;
                        double hs = c3r[0] * gr[0];
#pragma unroll
                        for (int k = 1; k < W; ++k) hs = fma(c3r[k], gr[k], hs);
                        H[r * P2_F3 + tx] = hs;
                    }
                }
            }
            __syncthreads();
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const double* const hp = H + o2r[q] * P2_F3 + tx;
                const double* const cr = c2s + (ty + 4 * q) * W;
                double os = cr[0] * hp[0];
#pragma unroll
                for (int k = 1; k < W; ++k) os = fma(cr[k], hp[k * P2_F3], os);
                g[W - 1][q] = os;
            }
            ++cb;
        }
        double w1[W];
#pragma unroll
        for (int k = 0; k < W; ++k) w1[k] = c1s[(j - j_lo) * W + k];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            double v = xa[q];
#pragma unroll
            for (int k = 0; k < W; ++k) v = fma(w1[k], g[k][q], v);
            if (okmask >> q & 1u) xp[q * rstep] = v;
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) xa[q] = xb[q];
        load_x(xb, xp + 2 * a.pldf, acc_in && j + 2 < j_hi);
        xp += a.pldf;
    }
}

// ---------------------------------------------------------------------------------------------
// Restriction.  CTA: 8 x 32 coarse tile of axes (2, 3) (one point per thread), marching along the
// FINE planes of a chunk of coarse planes.  Every thread owns NPT points of the halo'd fine tile:
//   per fine plane j:            acc[s] += R1[base(j)+s][j - start] * fine[j]  (registers, scatter form)
//   per completed coarse plane:  F = acc[0] (registers -> shared), H = F restricted along axis 3,
//                                out = H restricted along axis 2 -> global
// The first version kept the bookkeeping of the open coarse rows in the march (4 x [compare, branch,
// weight look-up] per plane, predicated loads, register copies of the prefetch pair): ncu at 515^3
// (profiles/r02_ncu_transfer_v2_first.txt) showed 2.4 warp instructions per fine point with the issue
// slots 47 % busy at two CTAs per SM -- instruction bound, 0.665 ms.  This version (0.456 ms) has each
// CTA first tabulate, per fine plane j of its chunk, the first open coarse row
// base(j), the weights of j in rows base(j) .. base(j)+NS-1 (zero where j is outside the row) and the
// number of rows that end at j; the march is then branch-free: NS broadcast weights x NPT FMAs per
// plane, unpredicated loads (dead slots re-read the tile origin; only their shared-memory store is masked), and the
// prefetch registers alternate by unrolling the plane loop twice instead of being copied.
constexpr int R2_C2 = 8, R2_C3 = 32, R2_RF2 = 22, R2_RF3 = 72, R2_MAXCH = 64;
constexpr int R2_FS = R2_RF3 + 1;
constexpr int R2_MAXJ = 2 * R2_MAXCH + 8;

template <int W, int NPT, int NS>
__global__ void __launch_bounds__(256, 2) restrict3d_tab_kernel(const TR3 a) {
    __shared__ double F[(R2_RF2 + 1) * R2_FS];
    __shared__ double H[(R2_RF2 + 8) * R2_C3];
    __shared__ double c2s[R2_C2 * W];
    __shared__ double c1s[R2_MAXCH * W];
    __shared__ double wt[R2_MAXJ * NS];
    __shared__ int s1s[R2_MAXCH];
    __shared__ int nes[R2_MAXJ];
    const int tid = threadIdx.x;
    const int tx = tid & (R2_C3 - 1), ty = tid >> 5;
    const int c3_0 = blockIdx.x * R2_C3, c2_0 = blockIdx.y * R2_C2;
    const int i_lo = blockIdx.z * a.chunk, i_hi = min(a.n1c, i_lo + a.chunk);
    const int c3l = min(R2_C3, a.n3c - c3_0), c2l = min(R2_C2, a.n2c - c2_0);
    const int f2lo = a.s2[c2_0], f3lo = a.s3[c3_0];
    const int f2hi = min(a.n2f - 1, a.s2[c2_0 + c2l - 1] + W - 1);
    const int f3hi = min(a.n3f - 1, a.s3[c3_0 + c3l - 1] + W - 1);
    const int nf2 = f2hi - f2lo + 1, nf3 = f3hi - f3lo + 1;
    const bool v3 = tx < c3l, v2 = ty < c2l;
    const int ni = i_hi - i_lo;

    for (int t = tid; t < (R2_RF2 + 1) * R2_FS; t += 256) F[t] = 0.0;
    for (int t = tid; t < (R2_RF2 + 8) * R2_C3; t += 256) H[t] = 0.0;
    for (int t = tid; t < R2_C2 * W; t += 256) {
        const int r = t / W;
        c2s[t] = r < c2l ? a.c2[(int64_t)c2_0 * W + t] : 0.0;
    }
    for (int t = tid; t < ni * W; t += 256) c1s[t] = a.c1[(int64_t)i_lo * W + t];
    for (int t = tid; t < ni; t += 256) s1s[t] = a.s1[i_lo + t];

    double c3r[W];
#pragma unroll
    for (int k = 0; k < W; ++k) c3r[k] = v3 ? a.c3[(int64_t)(c3_0 + tx) * W + k] : 0.0;
    const int o3 = v3 ? a.s3[c3_0 + tx] - f3lo : 0;
    const int o2 = v2 ? a.s2[c2_0 + ty] - f2lo : 0;
    // this thread's points of the halo'd fine tile (row-major over its true extent nf2 x nf3); dead
    // slots re-read the tile origin (unpredicated loads) and are masked out of the shared-memory store
    int go[NPT], so[NPT];
    unsigned fmask = 0;
#pragma unroll
    for (int m = 0; m < NPT; ++m) {
        const int slot = tid + 256 * m;
        const int r = slot / nf3, c = slot - r * nf3;
        const bool ok = r < nf2;
        if (ok) fmask |= 1u << m;
        go[m] = ok ? (int)((int64_t)r * a.ldf) + c : 0;
        so[m] = ok ? r * R2_FS + c : 0;
    }
    __syncthreads();

    // (a slab plan may start a row before its block: those taps carry zero coefficients)
    const int j_lo = max(0, s1s[0]);
    const int j_hi = min(a.n1f - 1, s1s[ni - 1] + W - 1);
    for (int t = tid; t <= j_hi - j_lo; t += 256) {
        const int j = j_lo + t;
        int base = 0;
        while (base < ni && s1s[base] + W - 1 < j) ++base;
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            const int i = base + s;
            const int tt = i < ni ? j - s1s[i] : -1;
            wt[t * NS + s] = (tt >= 0 && tt < W) ? c1s[i * W + tt] : 0.0;
        }
        int ne = 0;
        if (j == j_hi) ne = ni - base;
        else
            while (base + ne < ni && s1s[base + ne] + W - 1 == j) ++ne;
        nes[t] = ne;
    }
    __syncthreads();

    const double* fp = a.src + (int64_t)j_lo * a.pldf + (int64_t)f2lo * a.ldf + f3lo;
    const int64_t pld2 = 2 * a.pldf;
    double* outp = a.dst + (int64_t)i_lo * a.pldc + (int64_t)(c2_0 + ty) * a.ldc + c3_0 + tx;
    double fa[NPT], fb[NPT];
#pragma unroll
    for (int m = 0; m < NPT; ++m) fa[m] = __ldcs(fp + go[m]);
    if (j_lo + 1 <= j_hi) {
#pragma unroll
        for (int m = 0; m < NPT; ++m) fb[m] = __ldcs(fp + a.pldf + go[m]);
    } else {
#pragma unroll
        for (int m = 0; m < NPT; ++m) fb[m] = 0.0;
    }
    double acc[NS][NPT];
#pragma unroll
    for (int s = 0; s < NS; ++s)
#pragma unroll
        for (int m = 0; m < NPT; ++m) acc[s][m] = 0.0;

    // one fine plane: scatter f into the open rows, refill f with the plane two ahead, emit the rows
    // that end here
    auto step = [&](const int j, double (&f)[NPT]) {
        const double* const w = wt + (j - j_lo) * NS;
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            const double ws = w[s];
#pragma unroll
            for (int m = 0; m < NPT; ++m) acc[s][m] = fma(ws, f[m], acc[s][m]);
        }
        if (j + 2 <= j_hi) {
#pragma unroll
            for (int m = 0; m < NPT; ++m) f[m] = __ldcs(fp + pld2 + go[m]);
        }
        fp += a.pldf;
        const int ne = nes[j - j_lo];
#pragma unroll 1
        for (int e = 0; e < ne; ++e) {
#pragma unroll
            for (int m = 0; m < NPT; ++m)
                if (fmask >> m & 1u) F[so[m]] = acc[0][m];
            __syncthreads();
            if (v3) {
#pragma unroll 1
                for (int r = ty; r < nf2; r += 8) {
                    const double* const fr = F + r * R2_FS + o3;
                    double hs = c3r[0] * fr[0];
#pragma unroll
                    for (int k = 1; k < W; ++k) hs = fma(c3r[k], fr[k], hs);
                    H[r * R2_C3 + tx] = hs;
                }
            }
            __syncthreads();
            if (v2 && v3) {
                const double* const hp = H + o2 * R2_C3 + tx;
                const double* const cr = c2s + ty * W;
                double v = cr[0] * hp[0];
#pragma unroll
                for (int k = 1; k < W; ++k) v = fma(cr[k], hp[k * R2_C3], v);
                *outp = v;
            }
            outp += a.pldc;
#pragma unroll
            for (int s = 0; s + 1 < NS; ++s)
#pragma unroll
                for (int m = 0; m < NPT; ++m) acc[s][m] = acc[s + 1][m];
#pragma unroll
            for (int m = 0; m < NPT; ++m) acc[NS - 1][m] = 0.0;
        }
    };
#pragma unroll 1
    for (int j = j_lo; j <= j_hi; j += 2) {
        step(j, fa);
        if (j + 1 <= j_hi) step(j + 1, fb);
    }
}

template <int W>
int launch_prolong(const TR3& a, dim3 grid, cudaStream_t st) {
    POMS_LAUNCH(prolong3d_v2_kernel<W>, grid, st, a);
    CHECK_LAUNCH("poms_prolong_3d_v2");
    return 0;
}
template <int W, int NS>
int launch_restrict(const TR3& a, int npt, dim3 grid, cudaStream_t st) {
    if (npt <= 5) POMS_LAUNCH((restrict3d_tab_kernel<W, 5, NS>), grid, st, a);
    else if (npt == 6) POMS_LAUNCH((restrict3d_tab_kernel<W, 6, NS>), grid, st, a);
    else POMS_LAUNCH((restrict3d_tab_kernel<W, 7, NS>), grid, st, a);
    CHECK_LAUNCH("poms_restrict_3d_v2");
    return 0;
}

}  // namespace

extern "C" int poms_prolong_3d_v2(const double* coarse, double* fine, int n1f, int n2f, int n3f,
                                  int64_t ldf, int64_t pldf, int n1c, int n2c, int n3c, int64_t ldc,
                                  int64_t pldc, const int32_t* s1, const double* c1, int W1,
                                  const int32_t* s2, const double* c2, int W2, const int32_t* s3,
                                  const double* c3, int W3, const int32_t* s2_host,
                                  const int32_t* s3_host, int accumulate, void* stream) {
    if (!coarse || !fine || !s1 || !c1 || !s2 || !c2 || !s3 || !c3 || !s2_host || !s3_host)
        return t_bad_arg(1, "null pointer");
    if (n1f < 1 || n2f < 1 || n3f < 1 || n1c < 1 || n2c < 1 || n3c < 1) return t_bad_arg(3, "empty grid");
    if (W1 != W2 || W1 != W3 || W1 < 2 || W1 > 6)
        return t_bad_arg(15, "v2 needs one row width 2..6 on every axis (use poms_prolong_3d)");
    if ((int64_t)(n2c + 1) * ldc >= (1ll << 31)) return t_bad_arg(11, "coarse plane too large");
    const int e2 = rows_extent(s2_host, n2f, W2, n2c, P2_F2), e3 = rows_extent(s3_host, n3f, W3, n3c, P2_F3);
    if (e2 < 1 || e2 > P2_RC2 || e3 < 1 || e3 > P2_RC3 || e2 * e3 > 256 * p2_ng(W1))
        return t_bad_arg(22, "rows do not fit the fused-transfer tile (use poms_axis_gather)");
    TR3 a;
    a.src = coarse; a.dst = fine;
    a.n1f = n1f; a.n2f = n2f; a.n3f = n3f; a.n1c = n1c; a.n2c = n2c; a.n3c = n3c;
    a.ldf = ldf; a.pldf = pldf; a.ldc = ldc; a.pldc = pldc;
    a.s1 = s1; a.s2 = s2; a.s3 = s3; a.c1 = c1; a.c2 = c2; a.c3 = c3;
    a.accumulate = accumulate;
    const int g3 = (n3f + P2_F3 - 1) / P2_F3, g2 = (n2f + P2_F2 - 1) / P2_F2;
    // every chunk prolongs W coarse planes before its first fine plane (2 fine planes per coarse
    // plane afterwards): long chunks, but enough CTAs for ~4 waves of 148 SMs x 3
    int64_t nch = (148 * 3 * 4 + (int64_t)g3 * g2 - 1) / ((int64_t)g3 * g2);
    int chunk = (int)((n1f + nch - 1) / nch);
    if (chunk < 16) chunk = 16;
    if (chunk > P2_MAXCH) chunk = P2_MAXCH;
    if (chunk > n1f) chunk = n1f;
    chunk = (n1f + (n1f + chunk - 1) / chunk - 1) / ((n1f + chunk - 1) / chunk);   // equal chunks
    a.chunk = chunk;
    dim3 grid(g3, g2, (n1f + chunk - 1) / chunk);
    if (grid.y > 65535 || grid.z > 65535) return t_bad_arg(4, "grid too large");
    cudaStream_t st = (cudaStream_t)stream;
    switch (W1) {
        case 2: return launch_prolong<2>(a, grid, st);
        case 3: return launch_prolong<3>(a, grid, st);
        case 4: return launch_prolong<4>(a, grid, st);
        case 5: return launch_prolong<5>(a, grid, st);
        default: return launch_prolong<6>(a, grid, st);
    }
}

extern "C" int poms_restrict_3d_v2(const double* fine, double* coarse, int n1f, int n2f, int n3f,
                                   int64_t ldf, int64_t pldf, int n1c, int n2c, int n3c, int64_t ldc,
                                   int64_t pldc, const int32_t* s1, const double* c1, int W1,
                                   const int32_t* s2, const double* c2, int W2, const int32_t* s3,
                                   const double* c3, int W3, const int32_t* s1_host,
                                   const int32_t* s2_host, const int32_t* s3_host, void* stream) {
    if (!coarse || !fine || !s1 || !c1 || !s2 || !c2 || !s3 || !c3 || !s1_host || !s2_host || !s3_host)
        return t_bad_arg(1, "null pointer");
    if (n1f < 1 || n2f < 1 || n3f < 1 || n1c < 1 || n2c < 1 || n3c < 1) return t_bad_arg(3, "empty grid");
    if (W1 != W2 || W1 != W3 || W1 < 3 || W1 > 7)
        return t_bad_arg(15, "v2 needs one row width 3..7 on every axis (use poms_restrict_3d)");
    if ((int64_t)(R2_RF2 + 1) * ldf >= (1ll << 31)) return t_bad_arg(6, "fine row pitch too large");
    const int e2 = rows_extent(s2_host, n2c, W2, n2f, R2_C2), e3 = rows_extent(s3_host, n3c, W3, n3f, R2_C3);
    const int e1 = rows_extent(s1_host, n1c, W1, n1f, n1c, true);
    if (e1 < 1 || e2 < 1 || e2 > R2_RF2 || e3 < 1 || e3 > R2_RF3)
        return t_bad_arg(22, "rows do not fit the fused-transfer tile (use poms_axis_gather)");
    const int npt = (e2 * e3 + 255) / 256;
    const int nopen = rows_open(s1_host, n1c, W1);
    if (npt > 7 || nopen > 6) return t_bad_arg(22, "rows do not fit the fused-transfer registers");
    TR3 a;
    a.src = fine; a.dst = coarse;
    a.n1f = n1f; a.n2f = n2f; a.n3f = n3f; a.n1c = n1c; a.n2c = n2c; a.n3c = n3c;
    a.ldf = ldf; a.pldf = pldf; a.ldc = ldc; a.pldc = pldc;
    a.s1 = s1; a.s2 = s2; a.s3 = s3; a.c1 = c1; a.c2 = c2; a.c3 = c3;
    a.accumulate = 0;
    const int g3 = (n3c + R2_C3 - 1) / R2_C3, g2 = (n2c + R2_C2 - 1) / R2_C2;
    // every chunk re-reads the W-2 fine planes it shares with its neighbour: >= 16 coarse planes
    int64_t nch = (148 * 2 * 4 + (int64_t)g3 * g2 - 1) / ((int64_t)g3 * g2);
    int chunk = (int)((n1c + nch - 1) / nch);
    if (chunk < 16) chunk = 16;
    if (chunk > R2_MAXCH) chunk = R2_MAXCH;
    if (chunk > n1c) chunk = n1c;
    chunk = (n1c + (n1c + chunk - 1) / chunk - 1) / ((n1c + chunk - 1) / chunk);   // equal chunks
    a.chunk = chunk;
    dim3 grid(g3, g2, (n1c + chunk - 1) / chunk);
    if (grid.y > 65535 || grid.z > 65535) return t_bad_arg(4, "grid too large");
    cudaStream_t st = (cudaStream_t)stream;
    // the fine planes of one chunk must fit the per-plane tables of the kernel
    if (rows_extent(s1_host, n1c, W1, n1f, chunk, true) > R2_MAXJ)
        return t_bad_arg(22, "rows do not fit the fused-transfer tables (use poms_axis_gather)");
#define POMS_RS_CASE(WW)                                                           \
    case WW:                                                                       \
        return nopen <= 4 ? launch_restrict<WW, 4>(a, npt, grid, st)               \
                          : launch_restrict<WW, 6>(a, npt, grid, st);
    switch (W1) {
        POMS_RS_CASE(3)
        POMS_RS_CASE(4)
        POMS_RS_CASE(5)
        POMS_RS_CASE(6)
        default:
            return nopen <= 4 ? launch_restrict<7, 4>(a, npt, grid, st) : launch_restrict<7, 6>(a, npt, grid, st);
    }
#undef POMS_RS_CASE
}
