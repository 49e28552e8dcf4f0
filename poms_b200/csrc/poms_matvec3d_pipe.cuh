// poms_matvec3d_pipe.cuh -- K1 (3-D), round-2 kernel: the TMA-staged Kronecker mat-vec of
// poms_matvec3d_tma.cuh with the two halves of a plane step run in ANTI-PHASE inside the CTA.
//
// Round 1 (kron_matvec3d_tma_kernel): per plane  [stage 1: axis-3 pass, LSU bound] -> barrier ->
// [stages 2+3: axes 2 and 1, fp64 bound].  All warps of a CTA sit in the same stage, so the shared
// memory pipe and the fp64 pipe are loaded one after the other (ncu: both ~50 %, profiles/r01_ncu_*).
//
// Here stage 1 works ONE PLANE AHEAD of stages 2+3: between two barriers every warp runs stage 1 of
// plane j+1 and stages 2+3 of plane j -- warps 0-3 in that order, warps 4-7 in the opposite order, so
// each SM sub-partition (warp id mod 4) always holds one warp of each kind and the two pipes are
// loaded together.  su/sv are double buffered as before; still one barrier per plane.
// The rhs / x values of the fused epilogues are prefetched (cp.async into per-thread slots) one whole
// plane step ahead, right after the previous output plane has been emitted.
// The ring slot freed by stage 1 of plane j is refilled with plane j+NST (three planes in flight).
#pragma once

template <int P, int FORM, int EPI>
__global__ void __launch_bounds__(256, POMS_MV3_BLOCKS(P, FORM))
kron_matvec3d_pipe_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ MV3T g) {
    using C = MV3TCfg<P>;
    constexpr int W = C::W, T3 = C::T3, E = C::E, T2 = C::T2, R2 = C::R2, C3 = C::C3,
                  NST = C::NST, SH = C::SH;
    constexpr bool TWO = (FORM == POMS_FORM_SUM);
    constexpr int NX = 2 * P + 2;  // inputs of one output-column pair
    const MV3& a = g.a;
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
    const int grp = wid >> 2;                           // 0: stage 1 first, 1: stages 2+3 first
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
    // ring: plane j lives in slot (j - start) % NST; the first NST planes are requested here
    if (tid == 0) {
#pragma unroll
        for (int d = 0; d < NST; ++d) {
            if (start + d < end && valid(start + d)) {
                mbar_expect_tx(mbar + d, R2 * C3 * 8);
                tma_load_3d(ring + (size_t)d * (C::STAGE_BYTES / 8), &tmap, i3_0 - P, i2_0 - P,
                            start + d + a.glo, mbar + d);
            }
        }
    }
    const int64_t poff = (int64_t)(i2_0 + ty * E) * a.ld + i3;
    unsigned okmask = 0;
#pragma unroll
    for (int e = 0; e < E; ++e)
        if (v3 && (i2_0 + ty * E + e) < a.n2) okmask |= 1u << e;
    const int pslot = (ty * E) * T3 + tx;
    const bool wlive = __ballot_sync(0xffffffffu, okmask != 0) != 0;
    constexpr bool NEED_B = (EPI != POMS_EPI_STORE);
    const bool need_b = NEED_B && a.b != nullptr;   // AXPY without b: y = omega*v (zero initial guess)
    const bool need_x = (EPI == POMS_EPI_STORE && a.dot_out) || EPI == POMS_EPI_JACOBI;
    const bool need_pf = need_b || need_x;
    int64_t eoff = (int64_t)c_lo * a.pld + poff;    // offset of the NEXT output plane to be emitted
    unsigned phase_bits = 0;
    int st1 = 0, par1 = 0;   // stage 1: ring slot and su/sv buffer of the plane it handles next
    int u = 0, par = 0;      // stages 2+3: rotation slot and su/sv buffer
    int rslot = 0;           // (thread 0) ring slot refilled next

    // epilogue operands of the output plane at `eoff` into this thread's private slots
    auto prefetch_epi = [&]() {
#pragma unroll
        for (int e = 0; e < E; ++e) {
            if (okmask & (1u << e)) {
                if (need_b) cp_async8(pfb + pslot + e * T3, a.b + eoff + (int64_t)e * a.ld);
                if (need_x) cp_async8(pfx + pslot + e * T3, a.x + eoff + (int64_t)e * a.ld);
            }
        }
        cp_async_commit();
    };

    // ---- stage 1 of input plane jp: band pass along axis 3 into su/sv[par1] -------------------
    auto stage1 = [&](const int jp, auto steady_tag) {
        constexpr bool STEADY = decltype(steady_tag)::value;
        if (STEADY || valid(jp)) {
            mbar_wait(mbar + st1, (phase_bits >> st1) & 1u);
            phase_bits ^= (1u << st1);
            double* const sub = su + par1 * C::SU_DOUBLES;
            double* const svb = sv + par1 * C::SU_DOUBLES;
            const double* const sx = ring + (size_t)st1 * (C::STAGE_BYTES / 8) + 2 * lane;
            if (toep3) {
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
            // fix-up: column pairs holding a non-Toeplitz (boundary) row of M3 / K3 (see round-1 kernel)
            for (int it = tid; it < r_hi * nb; it += 256) {
                const int r = it / nb, q = it - r * nb;
                const int pr = q < tl ? q : th + (q - tl);
                const int fa = i3_0 + 2 * pr, fb = fa + 1;
                const bool oka = fa >= 0 && fa < a.n3, okb = fb >= 0 && fb < a.n3;
                double xr[NX];
                const double2* src = reinterpret_cast<const double2*>(
                    ring + (size_t)st1 * (C::STAGE_BYTES / 8) + r * C3 + 2 * pr);
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
        st1 = (st1 + 1 == NST) ? 0 : st1 + 1;
        par1 ^= 1;
    };

    // ---- stages 2+3 of input plane j1 (from su/sv[par]) and the epilogue of plane j1 - P --------
    auto stage23 = [&](const int j1, auto steady_tag) {
        constexpr bool STEADY = decltype(steady_tag)::value;
        const bool have = STEADY ? true : valid(j1);
        const int i1 = j1 - P;
        const bool emit = STEADY ? true : (i1 >= c_lo && i1 < c_hi);
        if (wlive) {
            double ta[E], tb[E], vout[E];
#pragma unroll
            for (int e = 0; e < E; ++e) ta[e] = tb[e] = 0.0;
            if (have) {
                const double* const up = su + par * C::SU_DOUBLES + (ty * E) * T3 + tx;
                const double* const vp = sv + par * C::SU_DOUBLES + (ty * E) * T3 + tx;
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
            const bool toep1 = STEADY ? true : (have && (j1 - P >= g.lo1) && (j1 + P < g.hi1));
            if (toep1) {
                rot_scatter<W, E, TWO>(u, acc, ta, tb, *(const double(*)[W])(TWO ? g.t1k : g.t1m),
                                       *(const double(*)[W]) g.t1m, vout);
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
            if (emit) {
                if (NEED_B || need_x) cp_async_wait<0>();
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
            // operands of the NEXT output plane: a whole plane step of latency hiding.  The slots are
            // private to this thread and were read just above (program order).
            if ((NEED_B || need_x) && need_pf && i1 + 1 >= c_lo && i1 + 1 < c_hi) prefetch_epi();
        }
        u = (u + 1 == W) ? 0 : u + 1;
        par ^= 1;
    };

    // ring slot of the plane stage 1 has just released (all warps are past the barrier): plane + NST
    auto refill = [&](const int jdone) {
        if (tid == 0) {
            const int jn = jdone + NST;
            if (jn < end && valid(jn)) {
                mbar_expect_tx(mbar + rslot, R2 * C3 * 8);
                tma_load_3d(ring + (size_t)rslot * (C::STAGE_BYTES / 8), &tmap, i3_0 - P, i2_0 - P,
                            jn + a.glo, mbar + rslot);
            }
            rslot = (rslot + 1 == NST) ? 0 : rslot + 1;
        }
    };

    // steady range of j1: plane j1 valid, emitting and Toeplitz in axis 1; plane j1+1 valid and < end
    int s_lo = max(max(start, -a.glo), max(c_lo + P, g.lo1 + P));
    int s_hi = min(min(end - 1, a.n1 + a.ghi - 1), min(c_hi + P, g.hi1 - P));
    if (s_hi < s_lo) s_hi = s_lo = start;

    stage1(start, std::false_type{});
    __syncthreads();
    refill(start);
    int j1 = start;
    auto step = [&](auto steady_tag) {
#pragma unroll 1
        for (int h = 0; h < 2; ++h) {
            if (h == grp) {
                if (decltype(steady_tag)::value || j1 + 1 < end) stage1(j1 + 1, steady_tag);
            } else {
                stage23(j1, steady_tag);
            }
        }
        __syncthreads();
        refill(j1 + 1);
    };
#pragma unroll 1
    for (; j1 < s_lo; ++j1) step(std::false_type{});
#pragma unroll 1
    for (; j1 < s_hi; ++j1) step(std::true_type{});
#pragma unroll 1
    for (; j1 < end; ++j1) step(std::false_type{});
    if (a.dot_out) {
        const double tot = block_sum(dsum, red);
        const unsigned nblk = gridDim.x * gridDim.y * gridDim.z;
        const unsigned bid = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
        grid_sum_finish(tot, a.dot_out, a.ws, nblk, bid, red);
    }
}
