// poms_matvec3d_v3.cuh -- K1 (3-D), round-2 kernel (variant 1): the TMA-staged Kronecker mat-vec
// of poms_matvec3d_tma.cuh restructured around three measured losses of the round-1 kernel
// (profiles/r01_ncu_kron_matvec3d_tma_v6_512.txt, r01_dynamic_instruction_mix_matvec3d.txt):
//
//  1. barrier stall (1.6 warp-cycles per issue: one __syncthreads per plane).  Here the CTA-wide
//     barrier is SPLIT: the axis-3 pass (stage 1) works one plane ahead and signals an mbarrier
//     (one arrival per warp); the register-only work of the current plane -- the axis-1 partial
//     sums (stage 3) and the epilogue -- sits between a warp's arrival and its wait, so the skew
//     between the warps of a CTA is absorbed instead of being waited out:
//        wait(S[j]) ; stage 2 (j) ; stage 1 (j+1) ; arrive(S[j+1]) ; stage 3 (j) ; epilogue (j)
//  2. bookkeeping (150 of 465 instructions per warp and plane): the rhs / x values of the fused
//     epilogues no longer come from four 8-byte cp.async per thread with 64-bit addresses, but
//     from ONE TMA tile load per plane issued by thread 0 (box = the CTA's 16 x 64 output tile,
//     3-slot ring, two output planes ahead).
//  3. fp64 work: the interior rows of the 1-D mass and stiffness matrices are SYMMETRIC Toeplitz
//     rows, so the Kronecker-SUM form shares the sums x[i-k] + x[i+k] between the M and the K pass
//     of an axis: stage 1  14 -> 11 and stage 2  21 -> 18 fp64 instructions per point (the host
//     routes a sum form with non-symmetric interior rows to the round-1 kernel).
// Everything else (tile shapes, TMA ring of halo'd input planes, boundary fix-ups, ragged tiles,
// rotating register partial sums, deterministic reduction) is the round-1 design.
#pragma once

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}

template <int P>
struct MV3V3Cfg : MV3TCfg<P> {
    using B = MV3TCfg<P>;
    static constexpr int EW = B::T3 + 2 * B::SH;       // epilogue tile: even start column, even width
    static constexpr int ETILE_BYTES = ((B::T2 * EW * 8 + 127) / 128) * 128;
    static constexpr int NES = 3;                      // epilogue tile ring
    static constexpr size_t smem_bytes(bool two, int ntiles) {
        return (size_t)B::NST * B::STAGE_BYTES + (size_t)(two ? 4 : 2) * B::SU_DOUBLES * 8 +
               (size_t)2 * B::T2 * B::W * 8 + (size_t)ntiles * NES * ETILE_BYTES + 32 * 8 /*mbar*/ +
               32 * 8 /*red*/;
    }
};

template <int P, int FORM, int EPI>
__global__ void __launch_bounds__(256, POMS_MV3_BLOCKS(P, FORM))
kron_matvec3d_v3_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ CUtensorMap tmb,
                        const __grid_constant__ CUtensorMap tmx, const __grid_constant__ MV3T g) {
    using C = MV3V3Cfg<P>;
    constexpr int W = C::W, T3 = C::T3, E = C::E, T2 = C::T2, R2 = C::R2, C3 = C::C3, NST = C::NST,
                  SH = C::SH, EW = C::EW, NES = C::NES;
    constexpr bool TWO = (FORM == POMS_FORM_SUM);
    constexpr int NX = 2 * P + 2;  // inputs of one output-column pair
    constexpr int STAGE_D = C::STAGE_BYTES / 8, ETILE_D = C::ETILE_BYTES / 8;
    const MV3& a = g.a;
    constexpr bool NEED_B = (EPI != POMS_EPI_STORE);
    const bool need_b = NEED_B && a.b != nullptr;   // AXPY without b: y = omega*v (zero initial guess)
    // g.dot_add (AXPY epilogue only): the fused reduction is sum y_out * z instead of sum (omega v)^2,
    // with z delivered through the x-tile ring (tmx is then a map of z): s.r of the CG driver rides in
    // the last smoother pass (/root/reference/sources/solvers.py:117-118)
    const bool dotz = (EPI == POMS_EPI_AXPY) && g.dot_add != 0;
    const bool need_x = (EPI == POMS_EPI_STORE && a.dot_out) || EPI == POMS_EPI_JACOBI || dotz;
    const bool need_t = need_b || need_x;

    extern __shared__ __align__(1024) unsigned char smem_raw[];
    double* const ring = reinterpret_cast<double*>(smem_raw);
    double* const su = ring + (size_t)NST * STAGE_D;
    double* const sv = su + 2 * C::SU_DOUBLES;
    double* const c2m = su + (TWO ? 4 : 2) * C::SU_DOUBLES;
    double* const c2k = c2m + T2 * W;
    double* const eb = c2k + T2 * W;                             // rhs tiles (if need_b)
    double* const ex = eb + (need_b ? NES * ETILE_D : 0);       // x tiles   (if need_x)
    uint64_t* const mbar = reinterpret_cast<uint64_t*>(ex + (need_x ? NES * ETILE_D : 0));
    uint64_t* const rfull = mbar;            // [NST]  input plane landed (TMA)
    uint64_t* const sfull = mbar + NST;      // [2]    su/sv buffer written by all 8 warps
    uint64_t* const efull = mbar + NST + 2;  // [NES]  epilogue tile(s) landed (TMA)
    double* const red = reinterpret_cast<double*>(mbar + 32);

    const int tid = threadIdx.x;
    const int tx = tid & (T3 - 1), ty = tid / T3;      // stage 2/3 mapping: column tx, rows ty*E..
    const int lane = tid & 31, wid = tid >> 5;          // stage 1 mapping: column pair `lane`
    const int i3_0 = blockIdx.x * T3 - SH, i2_0 = blockIdx.y * T2;
    const int c_lo = blockIdx.z * a.chunk;
    const int c_hi = min(a.n1, c_lo + a.chunk);
    const int i3 = i3_0 + tx;
    const bool v3 = i3 >= 0 && i3 < a.n3;
    const bool toep2_cta = (i2_0 >= g.lo2) && (i2_0 + T2 <= g.hi2);
    const bool toep2 = (i2_0 + ty * E >= g.lo2) && (i2_0 + ty * E + E <= g.hi2);
    const int r_hi = min(R2, a.n2 - i2_0 + 2 * P);
    const int qb = min(T3 / 2, (a.n3 - i3_0 + 1) >> 1);
    const int tl = min(max((g.lo3 - i3_0 + 1) >> 1, 0), qb);
    const int th = min(max((g.hi3 - i3_0) >> 1, tl), qb);
    const int nb = tl + (qb - th);
    const bool toep3 = lane >= tl && lane < th;

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < NST; ++s) mbar_init(rfull + s, 1);
        mbar_init(sfull + 0, 8);
        mbar_init(sfull + 1, 8);
#pragma unroll
        for (int s = 0; s < NES; ++s) mbar_init(efull + s, 1);
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

    // input planes [jv0, jv1) exist (owned + ghost planes); outputs c_lo .. c_hi-1 are completed by
    // the iterations j = c_lo+P .. c_hi+P-1 (the last ones may have no input plane: drain)
    const int jv0 = max(c_lo - P, -a.glo), jv1 = min(c_hi + P, a.n1 + a.ghi), jend = c_hi + P;
    const unsigned etx = (need_b ? T2 * EW * 8 : 0) + (need_x ? T2 * EW * 8 : 0);
    __syncthreads();  // barriers initialised, c2m/c2k staged
    if (tid == 0) {
#pragma unroll
        for (int d = 0; d < NST; ++d) {
            if (jv0 + d < jv1) {
                mbar_expect_tx(rfull + d, R2 * C3 * 8);
                tma_load_3d(ring + (size_t)d * STAGE_D, &tmap, i3_0 - P, i2_0 - P, jv0 + d + a.glo, rfull + d);
            }
        }
        if (need_t) {
#pragma unroll
            for (int d = 0; d < NES - 1; ++d) {
                if (c_lo + d < c_hi) {
                    mbar_expect_tx(efull + d, etx);
                    if (need_b) tma_load_3d(eb + (size_t)d * ETILE_D, &tmb, i3_0 - SH, i2_0, c_lo + d, efull + d);
                    if (need_x) tma_load_3d(ex + (size_t)d * ETILE_D, &tmx, i3_0 - SH, i2_0, c_lo + d + a.glo, efull + d);
                }
            }
        }
    }
    const int64_t poff = (int64_t)(i2_0 + ty * E) * a.ld + i3;
    unsigned okmask = 0;
#pragma unroll
    for (int e = 0; e < E; ++e)
        if (v3 && (i2_0 + ty * E + e) < a.n2) okmask |= 1u << e;
    const int eslot0 = (ty * E) * EW + tx + SH;     // this thread's first point inside an epilogue tile
    const bool wlive = __ballot_sync(0xffffffffu, okmask != 0) != 0;
    double* yp = a.y + ((int64_t)c_lo * a.pld + poff);   // this thread's first point of the NEXT output plane

    // ---- stage 1: band pass along axis 3 of the plane in ring slot `slot` into su/sv buffer `buf`
    auto stage1 = [&](const int slot, const unsigned par, const int buf) {
        mbar_wait(rfull + slot, par);
        const double* const base = ring + (size_t)slot * STAGE_D;
        double* const sub = su + buf * C::SU_DOUBLES;
        double* const svb = sv + buf * C::SU_DOUBLES;
        if (toep3) {
            const double* const sx = base + 2 * lane;
#pragma unroll 1
            for (int r = wid; r < r_hi; r += 8) {
                double xr[NX];
                const double2* src = reinterpret_cast<const double2*>(sx + r * C3);
#pragma unroll
                for (int q = 0; q < NX / 2; ++q) {
                    const double2 v2 = src[q];
                    xr[2 * q] = v2.x;
                    xr[2 * q + 1] = v2.y;
                }
                double ua, ub, va = 0.0, vb = 0.0;
                if (TWO) {
                    // symmetric interior rows: the pair sums are shared by the M and the K pass
                    ua = g.t3m[P] * xr[P];
                    ub = g.t3m[P] * xr[P + 1];
                    va = g.t3k[P] * xr[P];
                    vb = g.t3k[P] * xr[P + 1];
#pragma unroll
                    for (int k = 1; k <= P; ++k) {
                        const double sa = xr[P - k] + xr[P + k];
                        const double sb = xr[P + 1 - k] + xr[P + 1 + k];
                        ua = fma(g.t3m[P + k], sa, ua);
                        va = fma(g.t3k[P + k], sa, va);
                        ub = fma(g.t3m[P + k], sb, ub);
                        vb = fma(g.t3k[P + k], sb, vb);
                    }
                } else {
                    ua = g.t3m[0] * xr[0];
                    ub = g.t3m[0] * xr[1];
#pragma unroll
                    for (int k = 1; k < W; ++k) {
                        ua = fma(g.t3m[k], xr[k], ua);
                        ub = fma(g.t3m[k], xr[k + 1], ub);
                    }
                }
                *reinterpret_cast<double2*>(sub + r * T3 + 2 * lane) = make_double2(ua, ub);
                if (TWO) *reinterpret_cast<double2*>(svb + r * T3 + 2 * lane) = make_double2(va, vb);
            }
        }
        // fix-up: column pairs holding a non-Toeplitz (boundary) row of M3 / K3 (<= p+1 pairs per
        // domain end), spread over the whole CTA; coefficients from global memory (L1 resident)
        for (int it = tid; it < r_hi * nb; it += 256) {
            const int r = it / nb, q = it - r * nb;
            const int pr = q < tl ? q : th + (q - tl);
            const int fa = i3_0 + 2 * pr, fb = fa + 1;
            const bool oka = fa >= 0 && fa < a.n3, okb = fb >= 0 && fb < a.n3;
            double xr[NX];
            const double2* src = reinterpret_cast<const double2*>(base + r * C3 + 2 * pr);
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
        __syncwarp();
        if (lane == 0) mbar_arrive(sfull + buf);
    };

    // ---- stage 2: band pass along axis 2 of su/sv buffer `buf` for this thread's E rows ----------
    auto stage2 = [&](const int buf, double (&ta)[E], double (&tb)[E]) {
        const double* const up = su + buf * C::SU_DOUBLES + (ty * E) * T3 + tx;
        const double* const vp = sv + buf * C::SU_DOUBLES + (ty * E) * T3 + tx;
        if (toep2) {
            if (TWO) {
                double w[E + 2 * P];
#pragma unroll
                for (int r = 0; r < E + 2 * P; ++r) w[r] = up[r * T3];
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    ta[e] = g.t2m[P] * w[e + P];
                    tb[e] = g.t2k[P] * w[e + P];
#pragma unroll
                    for (int k = 1; k <= P; ++k) {
                        const double s = w[e + P - k] + w[e + P + k];
                        ta[e] = fma(g.t2m[P + k], s, ta[e]);
                        tb[e] = fma(g.t2k[P + k], s, tb[e]);
                    }
                }
#pragma unroll
                for (int r = 0; r < E + 2 * P; ++r) w[r] = vp[r * T3];
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    tb[e] = fma(g.t2m[P], w[e + P], tb[e]);
#pragma unroll
                    for (int k = 1; k <= P; ++k) {
                        const double s = w[e + P - k] + w[e + P + k];
                        tb[e] = fma(g.t2m[P + k], s, tb[e]);
                    }
                }
            } else {
#pragma unroll
                for (int r = 0; r < E + 2 * P; ++r) {
                    const double uv = up[r * T3];
#pragma unroll
                    for (int e = 0; e < E; ++e) {
                        const int k = r - e;
                        if (k >= 0 && k < W) ta[e] = (k == 0) ? g.t2m[0] * uv : fma(g.t2m[k], uv, ta[e]);
                    }
                }
            }
        } else {
#pragma unroll
            for (int e = 0; e < E; ++e) ta[e] = tb[e] = 0.0;
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
    };

    // state of the march: ring slot of the plane stage 1 handles next (+ phase bits of the ring),
    // su/sv buffer counter, epilogue tile slot (+ phase bits)
    int s1slot = 0;
    unsigned rphase = 0, ephase = 0;
    int cnt = 0, eslot = 0;

    if (jv0 < jv1) {
        stage1(0, 0u, 0);
        rphase ^= 1u;
        s1slot = (NST > 1) ? 1 : 0;
    }

    // thread 0, once every warp is past stage 1 of plane j and past the epilogue of plane j-2:
    // refill the ring slot of plane j with plane j+NST, the tile slot of output plane j-P-2 with j-P+1
    auto refill = [&](const int j, const bool ring_ok, const bool tile_ok) {
        if (tid == 0) {
            if (ring_ok) {
                const int sl = (s1slot == 0) ? NST - 1 : s1slot - 1;      // slot of plane j
                mbar_expect_tx(rfull + sl, R2 * C3 * 8);
                tma_load_3d(ring + (size_t)sl * STAGE_D, &tmap, i3_0 - P, i2_0 - P, j + NST + a.glo, rfull + sl);
            }
            if (tile_ok) {
                const int ti = j - P + 1;                                  // output plane of the NEXT iteration
                const int sl = (eslot + 1 == NES) ? 0 : eslot + 1;
                mbar_expect_tx(efull + sl, etx);
                if (need_b) tma_load_3d(eb + (size_t)sl * ETILE_D, &tmb, i3_0 - SH, i2_0, ti, efull + sl);
                if (need_x) tma_load_3d(ex + (size_t)sl * ETILE_D, &tmx, i3_0 - SH, i2_0, ti + a.glo, efull + sl);
            }
        }
    };

    // epilogue of output plane i1 for one point
    auto emit_point = [&](const int e, const double v, const double bv, const double xv, const double dg1,
                          const double dg2) {
        if (EPI == POMS_EPI_STORE) {
            yp[(int64_t)e * a.ld] = v;
            if (need_x) dsum = fma(xv, v, dsum);
        } else if (EPI == POMS_EPI_RESID) {
            const double rr = bv - v;
            yp[(int64_t)e * a.ld] = rr;
            dsum = fma(rr, rr, dsum);
        } else if (EPI == POMS_EPI_AXPY) {
            const double w_ = a.omega * v;
            const double yo = need_b ? bv + w_ : w_;
            yp[(int64_t)e * a.ld] = yo;
            dsum = dotz ? fma(yo, xv, dsum) : fma(w_, w_, dsum);
        } else {
            const double dg = TWO ? dg1 * dA[e] + dg2 * dB[e] : dg1 * dA[e];
            const double dr = a.omega * (bv - v) / dg;
            yp[(int64_t)e * a.ld] = (EPI == POMS_EPI_JACOBI) ? xv + dr : dr;
            dsum = fma(dr, dr, dsum);
        }
    };

    // ---- general plane step (boundary / ragged tiles, first and last planes of a chunk) ----------
    auto plane = [&](const int j, auto steady_tag) {
        constexpr bool STEADY = decltype(steady_tag)::value;
        const bool have = STEADY ? true : (j < jv1);
        double ta[E], tb[E], vout[E];
#pragma unroll
        for (int e = 0; e < E; ++e) ta[e] = tb[e] = 0.0;
        if (have) {
            mbar_wait(sfull + (cnt & 1), (cnt >> 1) & 1);
        } else {
            __syncthreads();   // drain planes (end of the last chunk only): orders the tile slots
        }
        const int i1 = j - P;
        const bool emit = STEADY ? true : (i1 >= c_lo);
        refill(j, have && (STEADY || j + NST < jv1),
               need_t && (STEADY || (i1 + 1 >= c_lo + NES - 1 && i1 + 1 < c_hi)));
        if (have && wlive) stage2(cnt & 1, ta, tb);
        if (STEADY || j + 1 < jv1) {
            stage1(s1slot, (rphase >> s1slot) & 1u, (cnt + 1) & 1);
            rphase ^= (1u << s1slot);
            s1slot = (s1slot + 1 == NST) ? 0 : s1slot + 1;
        }
        if (wlive) {
            // ---- stage 3: sliding axis-1 partial sums ----
            const bool toep1 = STEADY ? true : (have && (j - P >= g.lo1) && (j + P < g.hi1));
            if (toep1) {
                shift_scatter<W, E, TWO>(acc, ta, tb, *(const double(*)[W])(TWO ? g.t1k : g.t1m),
                                         *(const double(*)[W]) g.t1m, vout);
            } else {
                double c1k[W], c1m[W];
#pragma unroll
                for (int k = 0; k < W; ++k) {
                    const int o1 = j + P - k;
                    const bool ok = have && o1 >= 0 && o1 < a.n1;
                    if (TWO) {
                        c1k[k] = ok ? __ldg(a.k1 + (int64_t)o1 * W + k) : 0.0;
                        c1m[k] = ok ? __ldg(a.m1 + (int64_t)o1 * W + k) : 0.0;
                    } else {
                        c1k[k] = ok ? __ldg(a.m1 + (int64_t)o1 * W + k) : 0.0;
                        c1m[k] = 0.0;
                    }
                }
                shift_scatter<W, E, TWO>(acc, ta, tb, c1k, c1m, vout);
            }
        }
        // ---- epilogue: output plane i1 = j - P ----
        if (emit) {
            if (need_t) mbar_wait(efull + eslot, (ephase >> eslot) & 1u);
            ephase ^= (1u << eslot);
            if (wlive) {
                const double* const ebt = eb + (size_t)eslot * ETILE_D + eslot0;
                const double* const ext = ex + (size_t)eslot * ETILE_D + eslot0;
                double dg1 = 0.0, dg2 = 0.0;
                if (EPI == POMS_EPI_JACOBI || EPI == POMS_EPI_DINV) {
                    dg1 = TWO ? __ldg(a.k1 + (int64_t)i1 * W + P) : __ldg(a.m1 + (int64_t)i1 * W + P);
                    dg2 = TWO ? __ldg(a.m1 + (int64_t)i1 * W + P) : 0.0;
                }
#pragma unroll
                for (int e = 0; e < E; ++e)
                    if (okmask & (1u << e))
                        emit_point(e, vout[e], need_b ? ebt[e * EW] : 0.0, need_x ? ext[e * EW] : 0.0, dg1, dg2);
            }
            yp += a.pld;
            eslot = (eslot + 1 == NES) ? 0 : eslot + 1;
        }
        ++cnt;
    };

    // ---- fast plane step: steady range of an INTERIOR tile (every row and column of the halo'd tile
    // is a Toeplitz row, every point is inside the domain).  Both waits first, then ONE straight-line
    // block -- stage 2 and stage 3 of plane j with stage 1 of plane j+1 -- so that the compiler can
    // overlap the shared-memory latency of one stage with the FMA chains of the others.
    auto plane_fast = [&](const int j) {
        const int buf = cnt & 1;
        mbar_wait(sfull + buf, (cnt >> 1) & 1);
        mbar_wait(rfull + s1slot, (rphase >> s1slot) & 1u);
        refill(j, true, need_t);
        double ta[E], tb[E], vout[E];
#pragma unroll
        for (int e = 0; e < E; ++e) tb[e] = 0.0;
        {
            const double* const up = su + buf * C::SU_DOUBLES + (ty * E) * T3 + tx;
            const double* const vp = sv + buf * C::SU_DOUBLES + (ty * E) * T3 + tx;
            double w[E + 2 * P];
#pragma unroll
            for (int r = 0; r < E + 2 * P; ++r) w[r] = up[r * T3];
            if (TWO) {
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    ta[e] = g.t2m[P] * w[e + P];
                    tb[e] = g.t2k[P] * w[e + P];
#pragma unroll
                    for (int k = 1; k <= P; ++k) {
                        const double sk = w[e + P - k] + w[e + P + k];
                        ta[e] = fma(g.t2m[P + k], sk, ta[e]);
                        tb[e] = fma(g.t2k[P + k], sk, tb[e]);
                    }
                }
#pragma unroll
                for (int r = 0; r < E + 2 * P; ++r) w[r] = vp[r * T3];
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    tb[e] = fma(g.t2m[P], w[e + P], tb[e]);
#pragma unroll
                    for (int k = 1; k <= P; ++k) {
                        const double sk = w[e + P - k] + w[e + P + k];
                        tb[e] = fma(g.t2m[P + k], sk, tb[e]);
                    }
                }
            } else {
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    ta[e] = g.t2m[0] * w[e];
#pragma unroll
                    for (int k = 1; k < W; ++k) ta[e] = fma(g.t2m[k], w[e + k], ta[e]);
                }
            }
        }
#ifndef POMS_V3_S3_SHADOW
#define POMS_V3_S3_SHADOW 0
#endif
#if !POMS_V3_S3_SHADOW
        shift_scatter<W, E, TWO>(acc, ta, tb, *(const double(*)[W])(TWO ? g.t1k : g.t1m),
                                 *(const double(*)[W]) g.t1m, vout);
#endif
        {
            const double* const sx = ring + (size_t)s1slot * STAGE_D + 2 * lane;
            double* const sub = su + (buf ^ 1) * C::SU_DOUBLES + 2 * lane;
            double* const svb = sv + (buf ^ 1) * C::SU_DOUBLES + 2 * lane;
            auto row1 = [&](const int r) {
                double xr[NX];
                const double2* src = reinterpret_cast<const double2*>(sx + r * C3);
#pragma unroll
                for (int q = 0; q < NX / 2; ++q) {
                    const double2 v2 = src[q];
                    xr[2 * q] = v2.x;
                    xr[2 * q + 1] = v2.y;
                }
                double ua, ub, va = 0.0, vb = 0.0;
                if (TWO) {
                    ua = g.t3m[P] * xr[P];
                    ub = g.t3m[P] * xr[P + 1];
                    va = g.t3k[P] * xr[P];
                    vb = g.t3k[P] * xr[P + 1];
#pragma unroll
                    for (int k = 1; k <= P; ++k) {
                        const double sa = xr[P - k] + xr[P + k];
                        const double sb = xr[P + 1 - k] + xr[P + 1 + k];
                        ua = fma(g.t3m[P + k], sa, ua);
                        va = fma(g.t3k[P + k], sa, va);
                        ub = fma(g.t3m[P + k], sb, ub);
                        vb = fma(g.t3k[P + k], sb, vb);
                    }
                } else {
                    ua = g.t3m[0] * xr[0];
                    ub = g.t3m[0] * xr[1];
#pragma unroll
                    for (int k = 1; k < W; ++k) {
                        ua = fma(g.t3m[k], xr[k], ua);
                        ub = fma(g.t3m[k], xr[k + 1], ub);
                    }
                }
                *reinterpret_cast<double2*>(sub + r * T3) = make_double2(ua, ub);
                if (TWO) *reinterpret_cast<double2*>(svb + r * T3) = make_double2(va, vb);
            };
            // (giving warp 0 -- which also issues the TMA refills -- fewer rows was measured: no gain,
            // 5.87 vs 5.82 ms per iteration; the refill is not what the other warps wait for)
            row1(wid);
            row1(wid + 8);
            if (wid + 16 < R2) row1(wid + 16);
            if (R2 > 24 && wid + 24 < R2) row1(wid + 24);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(sfull + (buf ^ 1));
        rphase ^= (1u << s1slot);
        s1slot = (s1slot + 1 == NST) ? 0 : s1slot + 1;
#if POMS_V3_S3_SHADOW
        // experiment: axis-1 sums in the shadow between this warp's arrival and its next wait
        shift_scatter<W, E, TWO>(acc, ta, tb, *(const double(*)[W])(TWO ? g.t1k : g.t1m),
                                 *(const double(*)[W]) g.t1m, vout);
#endif
        // ---- epilogue ----
        if (need_t) mbar_wait(efull + eslot, (ephase >> eslot) & 1u);
        ephase ^= (1u << eslot);
        {
            const double* const ebt = eb + (size_t)eslot * ETILE_D + eslot0;
            const double* const ext = ex + (size_t)eslot * ETILE_D + eslot0;
            double dg1 = 0.0, dg2 = 0.0;
            if (EPI == POMS_EPI_JACOBI || EPI == POMS_EPI_DINV) {
                const int i1 = j - P;
                dg1 = TWO ? __ldg(a.k1 + (int64_t)i1 * W + P) : __ldg(a.m1 + (int64_t)i1 * W + P);
                dg2 = TWO ? __ldg(a.m1 + (int64_t)i1 * W + P) : 0.0;
            }
#pragma unroll
            for (int e = 0; e < E; ++e)
                emit_point(e, vout[e], need_b ? ebt[e * EW] : 0.0, need_x ? ext[e * EW] : 0.0, dg1, dg2);
        }
        yp += a.pld;
        eslot = (eslot + 1 == NES) ? 0 : eslot + 1;
        ++cnt;
    };

    // steady range: input plane j and j+1 exist, plane j+NST loadable, output emitted with its
    // successor's tile loadable, Toeplitz in axis 1
    int s_lo = max(jv0, max(c_lo + P, g.lo1 + P));
    s_lo = max(s_lo, c_lo + NES - 2 + P);
    int s_hi = min(jv1 - NST, min(c_hi + P - 1, g.hi1 - P));
    if (s_hi < s_lo) s_hi = s_lo = jv0;
    // interior tile: no boundary rows / columns, no ragged edge
    const bool cta_fast = toep2_cta && nb == 0 && qb == T3 / 2 && r_hi == R2 && i3_0 >= 0 &&
                          i3_0 + T3 <= a.n3 && i2_0 + T2 <= a.n2;
    int j = jv0;
#pragma unroll 1
    for (; j < s_lo; ++j) plane(j, std::false_type{});
    if (cta_fast) {
#pragma unroll 1
        for (; j < s_hi; ++j) plane_fast(j);
    } else {
#pragma unroll 1
        for (; j < s_hi; ++j) plane(j, std::true_type{});
    }
#pragma unroll 1
    for (; j < jend; ++j) plane(j, std::false_type{});
    if (a.dot_out) {
        const double tot = block_sum(dsum, red);
        const unsigned nbk = gridDim.x * gridDim.y * gridDim.z;
        const unsigned bid = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
        grid_sum_finish(tot, a.dot_out, a.ws, nbk, bid, red);
    }
}
