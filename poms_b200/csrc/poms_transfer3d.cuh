// poms_transfer3d.cuh -- fused 3-D knot-insertion transfers (included by poms_kernels.cu).
//
// The per-axis row gathers (poms_axis_gather) move every intermediate array through HBM: restriction
// 2.9 GB and prolongation 4.0 GB at 515^3 -> 259^3.  These kernels apply the three 1-D operators in
// one pass per fine plane, so the traffic is the fine array (read, or read + write) plus the coarse
// one: 1.2 GB and 2.3 GB.
//
// Rows are in the gather format of poms_axis_gather: out[i] = sum_{t<W} coef[i*W+t] * in[start[i]+t]
// with nondecreasing `start`; taps that fall outside the input carry a zero coefficient.
#pragma once

// ---------------------------------------------------------------------------------------------
// x_f (+)= (P1 (x) P2 (x) P3) e_c.   CTA: 16 x 64 fine tile of axes (2,3), marching along axis 1.
//   G = sum_t P1[j1][t] * e_c[start1[j1]+t]      on the coarse tile     (global -> shared)
//   H = G prolonged along axis 3                  (coarse rows x 64)     (shared -> shared)
//   out = H prolonged along axis 2                (16 x 64)              (shared -> global, += x_f)
// ---------------------------------------------------------------------------------------------
struct PR3 {
    const double* ec;
    double* xf;
    int n1f, n2f, n3f, n1c, n2c, n3c;
    int64_t ldf, pldf, ldc, pldc;
    const int32_t *s1, *s2, *s3;
    const double *c1, *c2, *c3;
    int W1, W2, W3;
    int chunk, accumulate;
};
constexpr int PR_F2 = 16, PR_F3 = 64, PR_RC2 = 16, PR_RC3 = 40, PR_WMAX = 8, PR_MAXCH = 64;

__global__ void __launch_bounds__(256, 4) prolong3d_kernel(const PR3 a) {
    __shared__ double G[PR_RC2][PR_RC3 + 1];
    __shared__ double H[PR_RC2][PR_F3];
    __shared__ double c2s[PR_F2][PR_WMAX];
    __shared__ double c1s[PR_MAXCH][PR_WMAX];
    __shared__ int o2s[PR_F2];
    __shared__ int s1s[PR_MAXCH];
    const int tid = threadIdx.x;
    const int f3_0 = blockIdx.x * PR_F3, f2_0 = blockIdx.y * PR_F2;
    const int j_lo = blockIdx.z * a.chunk, j_hi = min(a.n1f, j_lo + a.chunk);
    const int f3l = min(PR_F3, a.n3f - f3_0), f2l = min(PR_F2, a.n2f - f2_0);
    // coarse tile feeding this fine tile
    const int c2lo = a.s2[f2_0], c3lo = a.s3[f3_0];
    const int c2hi = min(a.n2c - 1, a.s2[f2_0 + f2l - 1] + a.W2 - 1);
    const int c3hi = min(a.n3c - 1, a.s3[f3_0 + f3l - 1] + a.W3 - 1);
    const int nr2 = c2hi - c2lo + 1, nr3 = c3hi - c3lo + 1;
    // rows of this CTA in shared memory: axis 2 of the tile, axis 1 of the chunk (uniform reads; the
    // streaming fine-grid traffic would keep evicting them from L1)
    for (int t = tid; t < PR_F2 * PR_WMAX; t += 256) {
        const int r = t / PR_WMAX, k = t - r * PR_WMAX;
        c2s[r][k] = (r < f2l && k < a.W2) ? a.c2[(int64_t)(f2_0 + r) * a.W2 + k] : 0.0;
        if (k == 0) o2s[r] = r < f2l ? a.s2[f2_0 + r] - c2lo : 0;
    }
    for (int t = tid; t < (j_hi - j_lo) * PR_WMAX; t += 256) {
        const int r = t / PR_WMAX, k = t - r * PR_WMAX;
        c1s[r][k] = k < a.W1 ? a.c1[(int64_t)(j_lo + r) * a.W1 + k] : 0.0;
        if (k == 0) s1s[r] = a.s1[j_lo + r];
    }
    // axis-3 row of this thread's column
    const int tx = tid & (PR_F3 - 1), ty = tid >> 6;
    const int lane = tid & 31, wid = tid >> 5;
    const bool v3 = tx < f3l;
    double c3r[PR_WMAX];
    int o3 = 0;
#pragma unroll
    for (int k = 0; k < PR_WMAX; ++k)
        c3r[k] = (v3 && k < a.W3) ? a.c3[(int64_t)(f3_0 + tx) * a.W3 + k] : 0.0;
    if (v3) o3 = a.s3[f3_0 + tx] - c3lo;
    // this thread's points of the coarse tile: rows wid, wid+8; columns lane, lane+32
    const double* gsrc[4];
    bool gok[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int r = wid + 8 * (q >> 1), c = lane + 32 * (q & 1);
        gok[q] = r < nr2 && c < nr3;
        gsrc[q] = a.ec + (int64_t)(c2lo + (gok[q] ? r : 0)) * a.ldc + (c3lo + (gok[q] ? c : 0));
    }
    __syncthreads();
    for (int j1 = j_lo; j1 < j_hi; ++j1) {
        // the fine values this thread updates: requested first, consumed last
        double xo[PR_F2 / 4];
        double* const dst = a.xf + (int64_t)j1 * a.pldf + (int64_t)f2_0 * a.ldf + f3_0 + tx;
#pragma unroll
        for (int q = 0; q < PR_F2 / 4; ++q) {
            const int r = ty + 4 * q;
            xo[q] = (a.accumulate && v3 && r < f2l) ? __ldcs(dst + (int64_t)r * a.ldf) : 0.0;
        }
        // ---- G: axis-1 combination of the coarse planes, on the coarse tile (independent loads) ----
        // (plane index clamped on BOTH sides: the rows of a slab plan may start before the plane block,
        // dist.slab_transfer_plan; those taps carry zero coefficients.  The lower clamp was missing until
        // tests/test_kernel_host_emulation.py ran this kernel under AddressSanitizer)
        const int p0 = s1s[j1 - j_lo];
        double w1[PR_WMAX];
#pragma unroll
        for (int k = 0; k < PR_WMAX; ++k) w1[k] = c1s[j1 - j_lo][k];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            if (gok[q]) {
                double gv[PR_WMAX];
#pragma unroll
                for (int k = 0; k < PR_WMAX; ++k)
                    gv[k] = k < a.W1 ? __ldg(gsrc[q] + (int64_t)max(0, min(p0 + k, a.n1c - 1)) * a.pldc) : 0.0;
                double gsum = 0.0;
#pragma unroll
                for (int k = 0; k < PR_WMAX; ++k) gsum = fma(w1[k], gv[k], gsum);
                G[wid + 8 * (q >> 1)][lane + 32 * (q & 1)] = gsum;
            }
        }
        __syncthreads();
        // ---- H: axis 3 ----
        if (v3) {
            for (int r = ty; r < nr2; r += 4) {
                double hsum = 0.0;
#pragma unroll
                for (int k = 0; k < PR_WMAX; ++k)
                    if (k < a.W3 && o3 + k < nr3) hsum = fma(c3r[k], G[r][o3 + k], hsum);
                H[r][tx] = hsum;
            }
        }
        __syncthreads();
        // ---- out: axis 2, accumulate into x_f ----
        if (v3) {
#pragma unroll
            for (int q = 0; q < PR_F2 / 4; ++q) {
                const int r = ty + 4 * q;
                if (r < f2l) {
                    const int o2 = o2s[r];
                    double osum = xo[q];
#pragma unroll
                    for (int k = 0; k < PR_WMAX; ++k)
                        if (k < a.W2 && o2 + k < nr2) osum = fma(c2s[r][k], H[o2 + k][tx], osum);
                    dst[(int64_t)r * a.ldf] = osum;
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// r_c = (R1 (x) R2 (x) R3) r_f.   CTA: 8 x 32 coarse tile (one point per thread), marching along the
// FINE axis 1: every fine plane is restricted in-plane (axis 3, then axis 2, through shared memory)
// and scattered into the partial sums of the coarse planes whose row of R1 contains it (at most
// RS_NS rows are open at a time; completed planes are emitted in order).  The fine tiles arrive by
// cp.async into a double buffer, one plane ahead of the arithmetic.
// ---------------------------------------------------------------------------------------------
struct RS3 {
    const double* rf;
    double* rc;
    int n1f, n2f, n3f, n1c, n2c, n3c;
    int64_t ldf, pldf, ldc, pldc;
    const int32_t *s1, *s2, *s3;
    const double *c1, *c2, *c3;
    int W1, W2, W3;
    int chunk;
};
constexpr int RS_C2 = 8, RS_C3 = 32, RS_RF2 = 22, RS_RF3 = 72, RS_NS = 6, RS_WMAX = 8, RS_MAXCH = 64;

__global__ void __launch_bounds__(256, 4) restrict3d_kernel(const RS3 a) {
    __shared__ double F[2][RS_RF2][RS_RF3 + 1];
    __shared__ double H[RS_RF2][RS_C3 + 1];
    __shared__ double c2s[RS_C2][RS_WMAX];
    __shared__ double c1s[RS_MAXCH][RS_WMAX];
    __shared__ int s1s[RS_MAXCH];
    const int tid = threadIdx.x;
    const int c3_0 = blockIdx.x * RS_C3, c2_0 = blockIdx.y * RS_C2;
    const int i_lo = blockIdx.z * a.chunk, i_hi = min(a.n1c, i_lo + a.chunk);
    const int c3l = min(RS_C3, a.n3c - c3_0), c2l = min(RS_C2, a.n2c - c2_0);
    const int f2lo = a.s2[c2_0], f3lo = a.s3[c3_0];
    const int f2hi = min(a.n2f - 1, a.s2[c2_0 + c2l - 1] + a.W2 - 1);
    const int f3hi = min(a.n3f - 1, a.s3[c3_0 + c3l - 1] + a.W3 - 1);
    const int nf2 = f2hi - f2lo + 1, nf3 = f3hi - f3lo + 1;
    const int tx = tid & (RS_C3 - 1), ty = tid >> 5;     // coarse column tx, coarse row ty
    const bool v3 = tx < c3l, v2 = ty < c2l;
    for (int t = tid; t < RS_C2 * RS_WMAX; t += 256) {
        const int r = t / RS_WMAX, k = t - r * RS_WMAX;
        c2s[r][k] = (r < c2l && k < a.W2) ? a.c2[(int64_t)(c2_0 + r) * a.W2 + k] : 0.0;
    }
    for (int t = tid; t < (i_hi - i_lo) * RS_WMAX; t += 256) {
        const int r = t / RS_WMAX, k = t - r * RS_WMAX;
        c1s[r][k] = k < a.W1 ? a.c1[(int64_t)(i_lo + r) * a.W1 + k] : 0.0;
        if (k == 0) s1s[r] = a.s1[i_lo + r];
    }
    double c3r[RS_WMAX];
    int o3 = 0, o2 = 0;
#pragma unroll
    for (int k = 0; k < RS_WMAX; ++k)
        c3r[k] = (v3 && k < a.W3) ? a.c3[(int64_t)(c3_0 + tx) * a.W3 + k] : 0.0;
    if (v3) o3 = a.s3[c3_0 + tx] - f3lo;
    if (v2) o2 = a.s2[c2_0 + ty] - f2lo;
    double acc[RS_NS];
#pragma unroll
    for (int s = 0; s < RS_NS; ++s) acc[s] = 0.0;
    int i_cur = i_lo;
    double* const out = a.rc + (int64_t)(c2_0 + ty) * a.ldc + c3_0 + tx;
    const int j_lo = a.s1[i_lo];
    const int j_hi = min(a.n1f - 1, a.s1[i_hi - 1] + a.W1 - 1);
    const double* const tile0 = a.rf + (int64_t)f2lo * a.ldf + f3lo;
    auto fetch = [&](int j1, int buf) {
        const double* src = tile0 + (int64_t)j1 * a.pldf;
        for (int r = ty; r < nf2; r += 8)
            for (int c = tx; c < nf3; c += 32) cp_async8(&F[buf][r][c], src + (int64_t)r * a.ldf + c);
    };
    fetch(j_lo, 0);
    cp_async_commit();
    __syncthreads();   // rows staged
    int buf = 0;
    for (int j1 = j_lo; j1 <= j_hi; ++j1) {
        if (j1 < j_hi) fetch(j1 + 1, buf ^ 1);
        cp_async_commit();
        cp_async_wait<1>();   // plane j1 has landed (all groups but the newest)
        __syncthreads();
        // ---- H: axis 3 (fine rows x coarse columns) ----
        if (v3) {
            for (int r = ty; r < nf2; r += 8) {
                double hsum = 0.0;
#pragma unroll
                for (int k = 0; k < RS_WMAX; ++k)
                    if (k < a.W3 && o3 + k < nf3) hsum = fma(c3r[k], F[buf][r][o3 + k], hsum);
                H[r][tx] = hsum;
            }
        }
        __syncthreads();
        // ---- axis 2, then scatter along axis 1 ----
        double v = 0.0;
        if (v2 && v3) {
#pragma unroll
            for (int k = 0; k < RS_WMAX; ++k)
                if (k < a.W2 && o2 + k < nf2) v = fma(c2s[ty][k], H[o2 + k][tx], v);
        }
#pragma unroll
        for (int s = 0; s < RS_NS; ++s) {
            const int i = i_cur + s;
            if (i < i_hi) {
                const int t = j1 - s1s[i - i_lo];
                if (t >= 0 && t < a.W1) acc[s] = fma(c1s[i - i_lo][t], v, acc[s]);
            }
        }
        while (i_cur < i_hi && (s1s[i_cur - i_lo] + a.W1 - 1 <= j1 || j1 == j_hi)) {
            if (v2 && v3) out[(int64_t)i_cur * a.pldc] = acc[0];
#pragma unroll
            for (int s = 0; s + 1 < RS_NS; ++s) acc[s] = acc[s + 1];
            acc[RS_NS - 1] = 0.0;
            ++i_cur;
        }
        buf ^= 1;
    }
}

// ---- host side: the tile bounds above are static, so the rows are checked against them ----------
static bool rows_fit(const int32_t* s, int n_out, int W, int n_in, int tile, int max_ext) {
    for (int i0 = 0; i0 < n_out; i0 += tile) {
        const int i1 = (i0 + tile < n_out ? i0 + tile : n_out) - 1;
        int hi = s[i1] + W - 1;
        if (hi > n_in - 1) hi = n_in - 1;
        if (hi - s[i0] + 1 > max_ext) return false;
    }
    for (int i = 1; i < n_out; ++i)
        if (s[i] < s[i - 1]) return false;
    return true;
}
static int rows_open_max(const int32_t* s, int n_out, int W) {
    // largest number of rows whose tap range [s, s+W-1] contains the same input index
    int best = 0, lo = 0;
    for (int i = 0; i < n_out; ++i) {
        while (s[lo] + W - 1 < s[i]) ++lo;
        if (i - lo + 1 > best) best = i - lo + 1;
    }
    return best;
}

extern "C" int poms_prolong_3d(const double* coarse, double* fine, int n1f, int n2f, int n3f, int64_t ldf,
                               int64_t pldf, int n1c, int n2c, int n3c, int64_t ldc, int64_t pldc,
                               const int32_t* s1, const double* c1, int W1, const int32_t* s2,
                               const double* c2, int W2, const int32_t* s3, const double* c3, int W3,
                               const int32_t* s2_host, const int32_t* s3_host, int accumulate,
                               void* stream) {
    if (!coarse || !fine || !s1 || !c1 || !s2 || !c2 || !s3 || !c3 || !s2_host || !s3_host)
        return bad_arg(1, "null pointer");
    if (n1f < 1 || n2f < 1 || n3f < 1 || n1c < 1 || n2c < 1 || n3c < 1) return bad_arg(3, "empty grid");
    if (W1 < 1 || W1 > PR_WMAX || W2 < 1 || W2 > PR_WMAX || W3 < 1 || W3 > PR_WMAX)
        return bad_arg(15, "row width must be 1..8");
    if (!rows_fit(s2_host, n2f, W2, n2c, PR_F2, PR_RC2) || !rows_fit(s3_host, n3f, W3, n3c, PR_F3, PR_RC3))
        return bad_arg(22, "rows do not fit the fused-transfer tile (use poms_axis_gather)");
    PR3 a;
    a.ec = coarse; a.xf = fine;
    a.n1f = n1f; a.n2f = n2f; a.n3f = n3f; a.n1c = n1c; a.n2c = n2c; a.n3c = n3c;
    a.ldf = ldf; a.pldf = pldf; a.ldc = ldc; a.pldc = pldc;
    a.s1 = s1; a.s2 = s2; a.s3 = s3; a.c1 = c1; a.c2 = c2; a.c3 = c3;
    a.W1 = W1; a.W2 = W2; a.W3 = W3;
    a.accumulate = accumulate;
    const int g3 = (n3f + PR_F3 - 1) / PR_F3, g2 = (n2f + PR_F2 - 1) / PR_F2;
    // ~4 waves of 148 SMs x 8 CTAs; at least 4 planes per CTA to amortise the per-CTA setup
    int64_t nch = (148 * 8 * 4 + (int64_t)g3 * g2 - 1) / ((int64_t)g3 * g2);
    int chunk = (int)((n1f + nch - 1) / nch);
    if (chunk < 4) chunk = 4;
    if (chunk > PR_MAXCH) chunk = PR_MAXCH;
    if (chunk > n1f) chunk = n1f;
    a.chunk = chunk;
    dim3 grid(g3, g2, (n1f + chunk - 1) / chunk);
    if (grid.y > 65535 || grid.z > 65535) return bad_arg(4, "grid too large");
    POMS_LAUNCH(prolong3d_kernel, grid, (cudaStream_t)stream, a);
    CHECK_LAUNCH("poms_prolong_3d");
    return 0;
}

extern "C" int poms_restrict_3d(const double* fine, double* coarse, int n1f, int n2f, int n3f, int64_t ldf,
                                int64_t pldf, int n1c, int n2c, int n3c, int64_t ldc, int64_t pldc,
                                const int32_t* s1, const double* c1, int W1, const int32_t* s2,
                                const double* c2, int W2, const int32_t* s3, const double* c3, int W3,
                                const int32_t* s1_host, const int32_t* s2_host, const int32_t* s3_host,
                                void* stream) {
    if (!coarse || !fine || !s1 || !c1 || !s2 || !c2 || !s3 || !c3 || !s1_host || !s2_host || !s3_host)
        return bad_arg(1, "null pointer");
    if (n1f < 1 || n2f < 1 || n3f < 1 || n1c < 1 || n2c < 1 || n3c < 1) return bad_arg(3, "empty grid");
    if (W1 < 1 || W1 > RS_WMAX || W2 < 1 || W2 > RS_WMAX || W3 < 1 || W3 > RS_WMAX)
        return bad_arg(15, "row width must be 1..8");
    if (!rows_fit(s2_host, n2c, W2, n2f, RS_C2, RS_RF2) || !rows_fit(s3_host, n3c, W3, n3f, RS_C3, RS_RF3) ||
        !rows_fit(s1_host, n1c, W1, n1f, n1c, n1f) || rows_open_max(s1_host, n1c, W1) > RS_NS)
        return bad_arg(22, "rows do not fit the fused-transfer tile (use poms_axis_gather)");
    RS3 a;
    a.rf = fine; a.rc = coarse;
    a.n1f = n1f; a.n2f = n2f; a.n3f = n3f; a.n1c = n1c; a.n2c = n2c; a.n3c = n3c;
    a.ldf = ldf; a.pldf = pldf; a.ldc = ldc; a.pldc = pldc;
    a.s1 = s1; a.s2 = s2; a.s3 = s3; a.c1 = c1; a.c2 = c2; a.c3 = c3;
    a.W1 = W1; a.W2 = W2; a.W3 = W3;
    const int g3 = (n3c + RS_C3 - 1) / RS_C3, g2 = (n2c + RS_C2 - 1) / RS_C2;
    int64_t nch = (148 * 8 * 4 + (int64_t)g3 * g2 - 1) / ((int64_t)g3 * g2);
    int chunk = (int)((n1c + nch - 1) / nch);
    // every chunk re-reads the W1-2 fine planes it shares with its neighbour: >= 8 coarse planes
    if (chunk < 8) chunk = 8;
    if (chunk > RS_MAXCH) chunk = RS_MAXCH;
    if (chunk > n1c) chunk = n1c;
    a.chunk = chunk;
    dim3 grid(g3, g2, (n1c + chunk - 1) / chunk);
    if (grid.y > 65535 || grid.z > 65535) return bad_arg(4, "grid too large");
    POMS_LAUNCH(restrict3d_kernel, grid, (cudaStream_t)stream, a);
    CHECK_LAUNCH("poms_restrict_3d");
    return 0;
}
