// poms_matvec3d_tma.cuh -- K1 (3-D) fast path: TMA-staged, mbarrier-pipelined Kronecker mat-vec.
// Included by poms_kernels.cu (single translation unit; uses its reduction / rot_scatter helpers).
//
// Per CTA: a 16 x 64 tile of axes (2,3), marching along axis 1.  For every input plane one elected
// thread issues ONE cp.async.bulk.tensor.3d (TMA) that lands the halo'd (16+2p) x (64+2p) tile in a
// shared-memory ring, zero-filled outside the domain by the hardware; an mbarrier signals arrival.
// The ring is 3 deep: two planes of prefetch are in flight ahead of the consumer (one plane was
// measured to leave the TMA latency exposed: 3.0 ms vs the generic kernel's 3.6 ms at 512^3; a
// fourth stage bought nothing and costs the second resident CTA).
// su/sv (axis-3 results) are double-buffered, so there is ONE __syncthreads per plane.
// Coefficients come from the kernel-parameter constant bank for the Toeplitz-interior rows of
// every axis (uniform knots: all but 2p rows per end).  The other rows: axis 3 -- the few boundary
// column pairs of a tile are recomputed by a CTA-wide fix-up from global memory; axis 2 -- a
// shared-memory table for the warps that hold boundary rows; axis 1 -- global memory on the few
// boundary planes.  Measured at 515^3, p = 3, three terms + residual: 1.10 ms.
#pragma once
#include <cuda.h>
#include <type_traits>

struct MV3T {
    MV3 a;
    double t1m[11], t1k[11], t2m[11], t2k[11], t3m[11], t3k[11];  // interior (Toeplitz) band rows
    int lo1, hi1, lo2, hi2, lo3, hi3;            // rows [lo, hi) of each axis equal to them
    int dot_add;
};

__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* tmap, int c0, int c1, int c2, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];\n" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)),
        "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"((unsigned)__cvta_generic_to_shared(bar))
        : "memory");
}

template <int P>
struct MV3TCfg {
    static constexpr int W = 2 * P + 1;
    static constexpr int T3 = 64, TY = 4, E = 4, T2 = TY * E;
    // TMA needs (innermost start coordinate * 8 B) 16-byte aligned (measured: an odd fp64
    // coordinate raises "illegal instruction").  Tiles therefore start at column 64*bx - (p&1):
    // the halo'd box then starts at the EVEN column 64*bx - (p&1) - p, and inside the tile the
    // inputs of the output-column pair (2l, 2l+1) are the 16-byte aligned columns 2l .. 2l+2p+1.
    static constexpr int SH = P & 1;
    static constexpr int R2 = T2 + 2 * P, C3 = T3 + 2 * P + 2 * SH;   // even
    static constexpr int PD = 2;        // prefetch distance (planes in flight ahead of the consumer)
    static constexpr int NST = PD + 1;  // ring depth
    static constexpr int STAGE_BYTES = ((R2 * C3 * 8 + 127) / 128) * 128;
    static constexpr int SU_DOUBLES = R2 * T3;
    static constexpr size_t smem_bytes(bool two) {
        return (size_t)NST * STAGE_BYTES + (size_t)(two ? 4 : 2) * SU_DOUBLES * 8 +
               (size_t)2 * T2 * W * 8 + 16 * 8 /*mbar*/ + 32 * 8 /*red*/ +
               (size_t)2 * T2 * T3 * 8 /* epilogue prefetch slots (b, x) */;
    }
};

POMS_HIDDEN int poms_mv3_tma_launch_p1(const CUtensorMap* tm3, const MV3T&, int form, int epi, int variant, int ntiles, dim3, cudaStream_t);
POMS_HIDDEN int poms_mv3_tma_launch_p2(const CUtensorMap* tm3, const MV3T&, int form, int epi, int variant, int ntiles, dim3, cudaStream_t);
POMS_HIDDEN int poms_mv3_tma_launch_p3(const CUtensorMap* tm3, const MV3T&, int form, int epi, int variant, int ntiles, dim3, cudaStream_t);
POMS_HIDDEN int poms_mv3_tma_launch_p4(const CUtensorMap* tm3, const MV3T&, int form, int epi, int variant, int ntiles, dim3, cudaStream_t);
POMS_HIDDEN int poms_mv3_tma_launch_p5(const CUtensorMap* tm3, const MV3T&, int form, int epi, int variant, int ntiles, dim3, cudaStream_t);

#if POMS_TU >= 1 && POMS_TU <= 5
#ifndef POMS_MV3_MINB
#define POMS_MV3_MINB 2
#endif
// two CTAs per SM whenever registers (2p+1 partial sums per point) and shared memory allow it
#define POMS_MV3_BLOCKS(P, FORM) (((P) <= 3 || ((P) == 4 && (FORM) == POMS_FORM_SINGLE)) ? POMS_MV3_MINB : 1)
template <int P, int FORM, int EPI>
__global__ void __launch_bounds__(256, POMS_MV3_BLOCKS(P, FORM))
kron_matvec3d_tma_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ MV3T g) {
    using C = MV3TCfg<P>;
    constexpr int W = C::W, T3 = C::T3, TY = C::TY, E = C::E, T2 = C::T2, R2 = C::R2, C3 = C::C3,
                  NST = C::NST, SH = C::SH, PD = C::PD;
    constexpr bool TWO = (FORM == POMS_FORM_SUM);
    constexpr int NX = 2 * P + 2;  // inputs of one output-column pair
    const MV3& a = g.a;
    // carve the dynamic shared memory with typed pointers (no integer round trip: the compiler
    // must see the shared address space, otherwise it emits generic LD/ST)
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    double* const ring = reinterpret_cast<double*>(smem_raw);
    double* const su = ring + (size_t)NST * (C::STAGE_BYTES / 8);
    double* const sv = su + 2 * C::SU_DOUBLES;
    double* const c2m = su + (TWO ? 4 : 2) * C::SU_DOUBLES;
    double* const c2k = c2m + T2 * W;
    uint64_t* const mbar = reinterpret_cast<uint64_t*>(c2k + T2 * W);
    double* const red = reinterpret_cast<double*>(mbar + 16);
    double* const pfb = red + 32;            // per-thread private slots: rhs of the output plane
    double* const pfx = pfb + T2 * T3;       //                           x at the output points

    const int tid = threadIdx.x;
    const int tx = tid & (T3 - 1), ty = tid / T3;      // stage 2/3 mapping: column tx, rows ty*E..
    const int lane = tid & 31, wid = tid >> 5;          // stage 1 mapping: column pair `lane`
    const int i3_0 = blockIdx.x * T3 - SH, i2_0 = blockIdx.y * T2;
    const int c_lo = blockIdx.z * a.chunk;
    const int c_hi = min(a.n1, c_lo + a.chunk);
    const int i3 = i3_0 + tx;
    const bool v3 = i3 >= 0 && i3 < a.n3;
    // axis-2 coefficients: Toeplitz per thread (warp-uniform: a warp holds ONE ty), table otherwise
    const bool toep2_cta = (i2_0 >= g.lo2) && (i2_0 + T2 <= g.hi2);
    const bool toep2 = (i2_0 + ty * E >= g.lo2) && (i2_0 + ty * E + E <= g.hi2);
    // rows of the halo'd tile that feed an output row inside the domain (ragged last tile row)
    const int r_hi = min(R2, a.n2 - i2_0 + 2 * P);
    // stage-1 columns of this lane and whether both are Toeplitz-interior rows of M3 / K3
    // pairs [0, qb) intersect the domain; pairs [tl, th) of them are Toeplitz, the other nb are not
    const int qb = min(T3 / 2, (a.n3 - i3_0 + 1) >> 1);
    const int tl = min(max((g.lo3 - i3_0 + 1) >> 1, 0), qb);
    const int th = min(max((g.hi3 - i3_0) >> 1, tl), qb);
    const int nb = tl + (qb - th);
    const bool toep3 = lane >= tl && lane < th;

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < NST; ++s) mbar_init(mbar + s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    }
    if (!toep2_cta) {
        for (int t = tid; t < T2 * W; t += blockDim.x) {
            const int r = t / W, k = t - r * W, i2 = i2_0 + r;
            c2m[t] = i2 < a.n2 ? a.m2[(int64_t)i2 * W + k] : 0.0;
            if (TWO) c2k[t] = i2 < a.n2 ? a.k2[(int64_t)i2 * W + k] : 0.0;
        }
    }
    double dA[E], dB[E];
    if (EPI == POMS_EPI_JACOBI || EPI == POMS_EPI_DINV) {
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const int i2 = i2_0 + ty * E + e;
            const bool ok = v3 && i2 < a.n2;
            const double m2d = ok ? a.m2[(int64_t)i2 * W + P] : 1.0;
            const double m3d = ok ? a.m3[(int64_t)i3 * W + P] : 1.0;
            dA[e] = m2d * m3d;
            dB[e] = 0.0;
            if (TWO) {
                const double k2d = ok ? a.k2[(int64_t)i2 * W + P] : 0.0;
                const double k3d = ok ? a.k3[(int64_t)i3 * W + P] : 0.0;
                dB[e] = k2d * m3d + m2d * k3d;
            }
        }
    }
    double acc[E][W];
#pragma unroll
    for (int e = 0; e < E; ++e)
#pragma unroll
        for (int k = 0; k < W; ++k) acc[e][k] = 0.0;
    double dsum = 0.0;

    const int start = c_lo - P, end = c_hi + P;
    auto valid = [&](int j1) { return j1 >= -a.glo && j1 < a.n1 + a.ghi; };
    __syncthreads();  // barriers initialised, c2m/c2k staged
    if (tid == 0) {
#pragma unroll
        for (int d = 0; d < PD; ++d) {
            if (start + d < end && valid(start + d)) {
                mbar_expect_tx(mbar + d, R2 * C3 * 8);
                tma_load_3d(ring + (size_t)d * (C::STAGE_BYTES / 8), &tmap, i3_0 - P, i2_0 - P,
                            start + d + a.glo, mbar + d);
            }
        }
    }
    // per-thread offset of its first output point inside a plane, validity mask of its E points
    const int64_t poff = (int64_t)(i2_0 + ty * E) * a.ld + i3;
    unsigned okmask = 0;
#pragma unroll
    for (int e = 0; e < E; ++e)
        if (v3 && (i2_0 + ty * E + e) < a.n2) okmask |= 1u << e;
    const int pslot = (ty * E) * T3 + tx;
    // warps whose 32 columns x E rows lie outside the domain (ragged last tiles) skip stages 2 and 3
    const bool wlive = __ballot_sync(0xffffffffu, okmask != 0) != 0;
    constexpr bool NEED_B = (EPI != POMS_EPI_STORE);
    const bool need_b = NEED_B && a.b != nullptr;   // AXPY without b: y = omega*v (zero initial guess)
    const bool need_x = (EPI == POMS_EPI_STORE && a.dot_out) || EPI == POMS_EPI_JACOBI;
    // output / rhs pointers of the NEXT plane to be emitted (planes are emitted in order)
    int64_t eoff = (int64_t)c_lo * a.pld + poff;
    unsigned phase_bits = 0;
    int u = 0, st = 0, t = 0;

    // One plane of the march.  STEADY = the plane is valid, completes an owned output plane, lies in
    // the Toeplitz-interior range of axis 1 and the prefetched plane exists: no checks are compiled.
    auto plane = [&](const int j1, auto steady_tag) {
        constexpr bool STEADY = decltype(steady_tag)::value;
        const bool have = STEADY ? true : valid(j1);
        const int i1 = j1 - P;
        const bool emit = STEADY ? true : (i1 >= c_lo && i1 < c_hi);
        if (NEED_B || need_x) {
            if (emit) {
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    if (okmask & (1u << e)) {
                        if (need_b) cp_async8(pfb + pslot + e * T3, a.b + eoff + (int64_t)e * a.ld);
                        if (need_x) cp_async8(pfx + pslot + e * T3, a.x + eoff + (int64_t)e * a.ld);
                    }
                }
            }
            cp_async_commit();
        }
        double* const sub = su + (t & 1) * C::SU_DOUBLES;
        double* const svb = sv + (t & 1) * C::SU_DOUBLES;
        if (have) {
            mbar_wait(mbar + st, (phase_bits >> st) & 1u);
            phase_bits ^= (1u << st);
            const double* const sx = ring + (size_t)st * (C::STAGE_BYTES / 8) + 2 * lane;
            // ---- stage 1: band pass along axis 3; lane = pair of output columns ----
            if (toep3) {
                // one row of the halo'd tile per warp and round; `row1` is the body for one row
                auto row1 = [&](const int r) {
                    double xr[NX];
                    const double2* src = reinterpret_cast<const double2*>(sx + r * C3);
#pragma unroll
                    for (int q = 0; q < NX / 2; ++q) {
                        const double2 v2 = src[q];
                        xr[2 * q] = v2.x;
                        xr[2 * q + 1] = v2.y;
                    }
                    double ua = g.t3m[0] * xr[0], ub = g.t3m[0] * xr[1];
                    double va = TWO ? g.t3k[0] * xr[0] : 0.0, vb = TWO ? g.t3k[0] * xr[1] : 0.0;
#pragma unroll
                    for (int k = 1; k < W; ++k) {
                        ua = fma(g.t3m[k], xr[k], ua);
                        ub = fma(g.t3m[k], xr[k + 1], ub);
                        if (TWO) {
                            va = fma(g.t3k[k], xr[k], va);
                            vb = fma(g.t3k[k], xr[k + 1], vb);
                        }
                    }
                    *reinterpret_cast<double2*>(sub + r * T3 + 2 * lane) = make_double2(ua, ub);
                    if (TWO) *reinterpret_cast<double2*>(svb + r * T3 + 2 * lane) = make_double2(va, vb);
                };
#pragma unroll 1
                for (int r = wid; r < r_hi; r += 8) row1(r);
            }
            // ---- stage 1 fix-up: column pairs holding a non-Toeplitz (boundary) row of M3 / K3.  They
            // are the first `tl` and the last `qb - th` pairs of the tile (<= p+1 pairs per domain end),
            // so the R2 x nb (row, pair) items are spread over the whole CTA: one round in a boundary
            // tile, nothing in an interior one; coefficients come from global memory (L1 resident).
            for (int it = tid; it < r_hi * nb; it += 256) {
                const int r = it / nb, q = it - r * nb;
                const int pr = q < tl ? q : th + (q - tl);
                const int fa = i3_0 + 2 * pr, fb = fa + 1;
                const bool oka = fa >= 0 && fa < a.n3, okb = fb >= 0 && fb < a.n3;
                double xr[NX];
                const double2* src = reinterpret_cast<const double2*>(
                    ring + (size_t)st * (C::STAGE_BYTES / 8) + r * C3 + 2 * pr);
#pragma unroll
                for (int q2 = 0; q2 < NX / 2; ++q2) {
                    const double2 v2 = src[q2];
                    xr[2 * q2] = v2.x;
                    xr[2 * q2 + 1] = v2.y;
                }
                double ua = 0.0, ub = 0.0, va = 0.0, vb = 0.0;
#pragma unroll
                for (int k = 0; k < W; ++k) {
                    const double ma = oka ? __ldg(a.m3 + (int64_t)fa * W + k) : 0.0;
                    const double mb = okb ? __ldg(a.m3 + (int64_t)fb * W + k) : 0.0;
                    ua = fma(ma, xr[k], ua);
                    ub = fma(mb, xr[k + 1], ub);
                    if (TWO) {
                        const double ka = oka ? __ldg(a.k3 + (int64_t)fa * W + k) : 0.0;
                        const double kb = okb ? __ldg(a.k3 + (int64_t)fb * W + k) : 0.0;
                        va = fma(ka, xr[k], va);
                        vb = fma(kb, xr[k + 1], vb);
                    }
                }
                *reinterpret_cast<double2*>(sub + r * T3 + 2 * pr) = make_double2(ua, ub);
                if (TWO) *reinterpret_cast<double2*>(svb + r * T3 + 2 * pr) = make_double2(va, vb);
            }
        }
        __syncthreads();
        // ---- producer: plane j1+PD into the slot of plane j1-1 (every thread is past its stage 1)
        if (tid == 0 && (STEADY || (j1 + PD < end && valid(j1 + PD)))) {
            const int sn = (st + PD) % NST;
            mbar_expect_tx(mbar + sn, R2 * C3 * 8);
            tma_load_3d(ring + (size_t)sn * (C::STAGE_BYTES / 8), &tmap, i3_0 - P, i2_0 - P,
                        j1 + PD + a.glo, mbar + sn);
        }
        double ta[E], tb[E], vout[E];
#pragma unroll
        for (int e = 0; e < E; ++e) ta[e] = tb[e] = 0.0;
        if (wlive) {
        if (have) {
            // ---- stage 2: band pass along axis 2, scatter form (no register window) ----
            const double* const up = sub + (ty * E) * T3 + tx;
            const double* const vp = svb + (ty * E) * T3 + tx;
            if (toep2) {
#pragma unroll
                for (int r = 0; r < E + 2 * P; ++r) {
                    const double uv = up[r * T3];
                    const double vv = TWO ? vp[r * T3] : 0.0;
#pragma unroll
                    for (int e = 0; e < E; ++e) {
                        const int k = r - e;
                        if (k >= 0 && k < W) {
                            ta[e] = (k == 0) ? g.t2m[0] * uv : fma(g.t2m[k], uv, ta[e]);
                            if (TWO) {
                                tb[e] = (k == 0) ? g.t2k[0] * uv : fma(g.t2k[k], uv, tb[e]);
                                tb[e] = fma(g.t2m[k], vv, tb[e]);
                            }
                        }
                    }
                }
            } else {
#pragma unroll
                for (int r = 0; r < E + 2 * P; ++r) {
                    const double uv = up[r * T3];
                    const double vv = TWO ? vp[r * T3] : 0.0;
#pragma unroll
                    for (int e = 0; e < E; ++e) {
                        const int k = r - e;
                        if (k >= 0 && k < W) {
                            const double cm = c2m[(ty * E + e) * W + k];
                            ta[e] = fma(cm, uv, ta[e]);
                            if (TWO) {
                                tb[e] = fma(c2k[(ty * E + e) * W + k], uv, tb[e]);
                                tb[e] = fma(cm, vv, tb[e]);
                            }
                        }
                    }
                }
            }
        }
        // ---- stage 3: rotating axis-1 partial sums ----
        const bool toep1 = STEADY ? true : (have && (j1 - P >= g.lo1) && (j1 + P < g.hi1));
        if (toep1) {
            rot_scatter<W, E, TWO>(u, acc, ta, tb, *(const double(*)[W])(TWO ? g.t1k : g.t1m), *(const double(*)[W]) g.t1m, vout);
        } else {
            double c1k[W], c1m[W];
#pragma unroll
            for (int k = 0; k < W; ++k) {
                const int o1 = j1 + P - k;
                const bool ok = have && o1 >= 0 && o1 < a.n1;
                if (TWO) {
                    c1k[k] = ok ? __ldg(a.k1 + (int64_t)o1 * W + k) : 0.0;
                    c1m[k] = ok ? __ldg(a.m1 + (int64_t)o1 * W + k) : 0.0;
                } else {
                    c1k[k] = ok ? __ldg(a.m1 + (int64_t)o1 * W + k) : 0.0;
                    c1m[k] = 0.0;
                }
            }
            rot_scatter<W, E, TWO>(u, acc, ta, tb, c1k, c1m, vout);
        }
        // ---- epilogue: output plane i1 = j1 - P ----
        if (NEED_B || need_x) cp_async_wait<0>();
        if (emit) {
            double dg1 = 0.0, dg2 = 0.0;
            if (EPI == POMS_EPI_JACOBI || EPI == POMS_EPI_DINV) {
                dg1 = TWO ? __ldg(a.k1 + (int64_t)i1 * W + P) : __ldg(a.m1 + (int64_t)i1 * W + P);
                dg2 = TWO ? __ldg(a.m1 + (int64_t)i1 * W + P) : 0.0;
            }
            double* const yp = a.y + eoff;
#pragma unroll
            for (int e = 0; e < E; ++e) {
                if (okmask & (1u << e)) {
                    const double v = vout[e];
                    if (EPI == POMS_EPI_STORE) {
                        yp[(int64_t)e * a.ld] = v;
                        if (need_x) dsum = fma(pfx[pslot + e * T3], v, dsum);
                    } else if (EPI == POMS_EPI_RESID) {
                        const double rr = pfb[pslot + e * T3] - v;
                        yp[(int64_t)e * a.ld] = rr;
                        dsum = fma(rr, rr, dsum);
                    } else if (EPI == POMS_EPI_AXPY) {
                        const double w_ = a.omega * v;
                        yp[(int64_t)e * a.ld] = need_b ? pfb[pslot + e * T3] + w_ : w_;
                        dsum = fma(w_, w_, dsum);
                    } else {
                        const double dg = TWO ? dg1 * dA[e] + dg2 * dB[e] : dg1 * dA[e];
                        const double dr = a.omega * (pfb[pslot + e * T3] - v) / dg;
                        yp[(int64_t)e * a.ld] = (EPI == POMS_EPI_JACOBI) ? pfx[pslot + e * T3] + dr : dr;
                        dsum = fma(dr, dr, dsum);
                    }
                }
            }
            eoff += a.pld;
        }
        }  // wlive
        u = (u + 1 == W) ? 0 : u + 1;
        st = (st + 1 == NST) ? 0 : st + 1;
        ++t;
    };

    // steady range of input planes: valid, emitting, Toeplitz in axis 1, and plane j1+PD loadable
    int s_lo = max(max(start, -a.glo), max(c_lo + P, g.lo1 + P));
    int s_hi = min(min(end, a.n1 + a.ghi), min(c_hi + P, g.hi1 - P));
    s_hi = min(s_hi, min(end, a.n1 + a.ghi) - PD);
    if (s_hi < s_lo) s_hi = s_lo = start;
    int j1 = start;
#pragma unroll 1
    for (; j1 < s_lo; ++j1) plane(j1, std::false_type{});
#pragma unroll 1
    for (; j1 < s_hi; ++j1) plane(j1, std::true_type{});
#pragma unroll 1
    for (; j1 < end; ++j1) plane(j1, std::false_type{});
    if (a.dot_out) {
        const double tot = block_sum(dsum, red);
        const unsigned nb = gridDim.x * gridDim.y * gridDim.z;
        const unsigned bid = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
        grid_sum_finish(tot, a.dot_out, a.ws, nb, bid, red);
    }
}

// ---- host side -----------------------------------------------------------------------------
#include "poms_matvec3d_v3.cuh"

// tm3: [0] halo'd input planes of x, [1] rhs tiles, [2] x tiles (variant 1 only; unused ones repeat [0])
template <int P, int FORM, int EPI, int VAR>
static int launch_mv3_tma_inst(const CUtensorMap* tm3, const MV3T& g, int ntiles, dim3 grid, cudaStream_t st) {
    if (VAR == 0) {
        const size_t smem = MV3TCfg<P>::smem_bytes(FORM == POMS_FORM_SUM);
        auto kern = kron_matvec3d_tma_kernel<P, FORM, EPI>;
        static bool attr_set = false;
        if (!attr_set) {
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return fail_cuda(e, "cudaFuncSetAttribute(tma)");
            attr_set = true;
        }
        kern<<<grid, 256, smem, st>>>(tm3[0], g);
    } else {
        auto kern = kron_matvec3d_v3_kernel<P, FORM, EPI>;
        static bool attr_set = false;
        if (!attr_set) {
            const size_t smax = MV3V3Cfg<P>::smem_bytes(FORM == POMS_FORM_SUM, 2);
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smax);
            if (e != cudaSuccess) return fail_cuda(e, "cudaFuncSetAttribute(v3)");
            attr_set = true;
        }
        const size_t smem = MV3V3Cfg<P>::smem_bytes(FORM == POMS_FORM_SUM, ntiles);
        kern<<<grid, 256, smem, st>>>(tm3[0], tm3[1], tm3[2], g);
    }
    return 0;
}
template <int P, int FORM, int VAR>
static int launch_mv3_tma_epi(const CUtensorMap* tm3, const MV3T& g, int epi, int ntiles, dim3 grid, cudaStream_t st) {
    switch (epi) {
        case POMS_EPI_STORE: return launch_mv3_tma_inst<P, FORM, POMS_EPI_STORE, VAR>(tm3, g, ntiles, grid, st);
        case POMS_EPI_RESID: return launch_mv3_tma_inst<P, FORM, POMS_EPI_RESID, VAR>(tm3, g, ntiles, grid, st);
        case POMS_EPI_JACOBI: return launch_mv3_tma_inst<P, FORM, POMS_EPI_JACOBI, VAR>(tm3, g, ntiles, grid, st);
        case POMS_EPI_DINV: return launch_mv3_tma_inst<P, FORM, POMS_EPI_DINV, VAR>(tm3, g, ntiles, grid, st);
        case POMS_EPI_AXPY: return launch_mv3_tma_inst<P, FORM, POMS_EPI_AXPY, VAR>(tm3, g, ntiles, grid, st);
        default: return bad_arg(19, "epilogue");
    }
}
#define POMS_CAT_(a, b) a##b
#define POMS_CAT(a, b) POMS_CAT_(a, b)
int POMS_CAT(poms_mv3_tma_launch_p, POMS_TU)(const CUtensorMap* tm3, const MV3T& g, int form, int epi, int variant,
                                             int ntiles, dim3 grid, cudaStream_t st) {
    constexpr int P = POMS_TU;
    if (variant == 0) {
        if (form == POMS_FORM_SINGLE) return launch_mv3_tma_epi<P, POMS_FORM_SINGLE, 0>(tm3, g, epi, ntiles, grid, st);
        return launch_mv3_tma_epi<P, POMS_FORM_SUM, 0>(tm3, g, epi, ntiles, grid, st);
    }
    if (form == POMS_FORM_SINGLE) return launch_mv3_tma_epi<P, POMS_FORM_SINGLE, 1>(tm3, g, epi, ntiles, grid, st);
    return launch_mv3_tma_epi<P, POMS_FORM_SUM, 1>(tm3, g, epi, ntiles, grid, st);
}
#endif  // POMS_TU in 1..5

#if POMS_TU == 0
#include <unordered_map>
#include <mutex>

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled get_encode_tiled() {
    static PFN_encodeTiled fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (PFN_encodeTiled)p;
    }
    return fn;
}

// Tensor maps are pure functions of (base address, extents, pitches, box): the V-cycle applies the
// same few operators to the same few vectors thousands of times, so the encoded descriptors are kept
// (round 1 re-encoded one per launch on the host, 576 times per solve).
struct TmapKey {
    uint64_t addr, n3, n2, n1, ld, pld, box;
    bool operator==(const TmapKey& o) const {
        return addr == o.addr && n3 == o.n3 && n2 == o.n2 && n1 == o.n1 && ld == o.ld && pld == o.pld && box == o.box;
    }
};
struct TmapKeyHash {
    size_t operator()(const TmapKey& k) const {
        uint64_t h = 1469598103934665603ull;
        const uint64_t v[7] = {k.addr, k.n3, k.n2, k.n1, k.ld, k.pld, k.box};
        for (int i = 0; i < 7; ++i) { h ^= v[i]; h *= 1099511628211ull; }
        return (size_t)h;
    }
};
static int get_tmap(PFN_encodeTiled enc, const double* base, int n3, int n2, int n1tot, int64_t ld, int64_t pld,
                    int boxw, int boxh, CUtensorMap* out) {
    static std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> cache;
    static std::mutex mu;
    const TmapKey key{(uint64_t)(uintptr_t)base, (uint64_t)n3, (uint64_t)n2, (uint64_t)n1tot, (uint64_t)ld,
                      (uint64_t)pld, ((uint64_t)boxw << 32) | (uint64_t)boxh};
    std::lock_guard<std::mutex> lk(mu);
    auto it = cache.find(key);
    if (it != cache.end()) { *out = it->second; return 0; }
    cuuint64_t dims[3] = {(cuuint64_t)n3, (cuuint64_t)n2, (cuuint64_t)n1tot};
    cuuint64_t strides[2] = {(cuuint64_t)ld * 8, (cuuint64_t)pld * 8};
    cuuint32_t box[3] = {(cuuint32_t)boxw, (cuuint32_t)boxh, 1};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, (void*)base, dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return 1;
    if (cache.size() > 4096) cache.clear();   // addresses of freed temporaries: bounded, not LRU
    cache.emplace(key, *out);
    return 0;
}

// kernel variant: 1 = split-barrier kernel of poms_matvec3d_v3.cuh (round 2, default), 0 = round-1 kernel (A/B timing, tests)
static int g_mv3_variant = -1;
extern "C" void poms_set_matvec3d_variant(int v) { g_mv3_variant = v; }
static int mv3_variant() {
    if (g_mv3_variant < 0) {
        const char* e = getenv("POMS_B200_MV3_VARIANT");
        g_mv3_variant = e ? atoi(e) : 1;
    }
    return g_mv3_variant;
}

// returns 0 on success, 1 if the TMA path does not apply (caller falls back to the generic kernel),
// other values are errors
static int try_matvec3d_tma(const MV3& a0, int p, int form, int epilogue, const double* toep, const int* toep_rng,
                            cudaStream_t st, const double* dot_with = nullptr, int* fused = nullptr) {
    if (fused) *fused = 0;
    if (((uintptr_t)a0.x & 15) || (a0.ld & 1) || (a0.pld & 1)) return 1;
    if (a0.n3 < 8 || a0.n2 < 4) return 1;  // tiny grids: generic kernel
    PFN_encodeTiled enc = get_encode_tiled();
    if (!enc) return 1;
    const int W = 2 * p + 1;
    MV3T g;
    g.a = a0;
    g.lo1 = g.hi1 = g.lo2 = g.hi2 = g.lo3 = g.hi3 = 0;
    g.dot_add = 0;
    for (int k = 0; k < 11; ++k) g.t1m[k] = g.t1k[k] = g.t2m[k] = g.t2k[k] = g.t3m[k] = g.t3k[k] = 0.0;
    if (toep && toep_rng) {
        // toep: [axis][m|k][W] for axes 1, 2, 3 ; toep_rng: [axis][lo, hi)
        for (int k = 0; k < W; ++k) {
            g.t1m[k] = toep[(0 * 2 + 0) * W + k];
            g.t1k[k] = toep[(0 * 2 + 1) * W + k];
            g.t2m[k] = toep[(1 * 2 + 0) * W + k];
            g.t2k[k] = toep[(1 * 2 + 1) * W + k];
            g.t3m[k] = toep[(2 * 2 + 0) * W + k];
            g.t3k[k] = toep[(2 * 2 + 1) * W + k];
        }
        g.lo1 = toep_rng[0];
        g.hi1 = toep_rng[1];
        g.lo2 = toep_rng[2];
        g.hi2 = toep_rng[3];
        g.lo3 = toep_rng[4];
        g.hi3 = toep_rng[5];
    }
    const int sh = p & 1;
    const int g3 = (a0.n3 + sh + 63) / 64, g2 = (a0.n2 + 15) / 16;
    g.a.chunk = pick_chunk(a0.n1, (int64_t)g3 * g2, p);
    const int g1 = (a0.n1 + g.a.chunk - 1) / g.a.chunk;
    if ((int64_t)g3 * g2 * g1 > POMS_MAX_PARTIALS) return 1;
    CUtensorMap tm3[3];
    if (get_tmap(enc, a0.x - (int64_t)a0.glo * a0.pld, a0.n3, a0.n2, a0.n1 + a0.glo + a0.ghi, a0.ld, a0.pld,
                 64 + 2 * p + 2 * sh, 16 + 2 * p, &tm3[0]))
        return 1;
    tm3[1] = tm3[0];
    tm3[2] = tm3[0];
    int var = mv3_variant();
    int ntiles = 0;
    if (var != 0) {
        // variant 1 (poms_matvec3d_v3.cuh): the sum form shares the pair sums of the SYMMETRIC interior
        // rows between its M and K passes; the epilogue operands arrive as TMA tiles (16-byte aligned)
        // (assembled interior rows are symmetric to rounding only: the kernel reads the upper half
        // of a row for both sides, a relative change of the coefficients below 1e-13)
        bool sym = true;
        if (form == POMS_FORM_SUM) {
            const double* rows[4] = {g.t2m, g.t2k, g.t3m, g.t3k};
            for (int r = 0; r < 4; ++r) {
                double mx = 0.0;
                for (int k = 0; k < W; ++k) mx = fmax(mx, fabs(rows[r][k]));
                for (int k = 0; k < W; ++k) sym = sym && fabs(rows[r][k] - rows[r][W - 1 - k]) <= 1e-13 * mx;
            }
        }
        const bool need_b = epilogue != POMS_EPI_STORE && a0.b != nullptr;
        const bool dotz = epilogue == POMS_EPI_AXPY && dot_with && a0.dot_out && !((uintptr_t)dot_with & 15);
        const bool need_x = (epilogue == POMS_EPI_STORE && a0.dot_out) || epilogue == POMS_EPI_JACOBI || dotz;
        if (!sym || (need_b && ((uintptr_t)a0.b & 15))) {
            var = 0;
        } else {
            // the pair-sum form uses the MEAN of the two halves of a row, so that row sums (K 1 = 0,
            // partition of unity) carry no systematic bias from the rounding-level asymmetry
            if (form == POMS_FORM_SUM) {
                double* rows[4] = {g.t2m, g.t2k, g.t3m, g.t3k};
                for (int r = 0; r < 4; ++r)
                    for (int k = 1; k <= p; ++k)
                        rows[r][p + k] = rows[r][p - k] = 0.5 * (rows[r][p + k] + rows[r][p - k]);
            }
            if (need_b) {
                if (get_tmap(enc, a0.b, a0.n3, a0.n2, a0.n1, a0.ld, a0.pld, 64 + 2 * sh, 16, &tm3[1])) return 1;
                ++ntiles;
            }
            if (need_x) {
                const double* xt = dotz ? dot_with : a0.x;       // same slab layout (ghost planes) as x
                if (get_tmap(enc, xt - (int64_t)a0.glo * a0.pld, a0.n3, a0.n2, a0.n1 + a0.glo + a0.ghi, a0.ld,
                             a0.pld, 64 + 2 * sh, 16, &tm3[2]))
                    return 1;
                ++ntiles;
            }
            if (dotz) {
                g.dot_add = 1;
                if (fused) *fused = 1;
            }
        }
    }
    dim3 grid(g3, g2, g1);
    switch (p) {
        case 1: return poms_mv3_tma_launch_p1(tm3, g, form, epilogue, var, ntiles, grid, st);
        case 2: return poms_mv3_tma_launch_p2(tm3, g, form, epilogue, var, ntiles, grid, st);
        case 3: return poms_mv3_tma_launch_p3(tm3, g, form, epilogue, var, ntiles, grid, st);
        case 4: return poms_mv3_tma_launch_p4(tm3, g, form, epilogue, var, ntiles, grid, st);
        case 5: return poms_mv3_tma_launch_p5(tm3, g, form, epilogue, var, ntiles, grid, st);
        default: return bad_arg(11, "p must be 1..5");
    }
}
#endif  // POMS_TU == 0
