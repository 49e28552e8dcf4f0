// poms_matvec2d_tma.cuh -- K1 (2-D), round-2 kernel: warp-autonomous, TMA-staged Kronecker mat-vec
// Y = A1 X A2^T (or the Kronecker sum K1 (x) M2 + M1 (x) K2) for /root/reference/sources/
// kron_product.py:56-89 and the operator mat-vec of the 2-D solvers (sources/solvers.py:85,103,209).
//
// The round-1 2-D kernel staged eight rows per CTA with synchronous loads and two CTA barriers, and
// kept every thread's 2 x (2p+1) in-row coefficients in registers (250 registers at p = 5).  In two
// dimensions no stage needs another warp's results -- the in-row band pass (axis 2) reads the input
// row, and the cross-row pass (axis 1) is a sliding window of partial sums in the SAME thread's
// registers -- so here every WARP is a pipeline of its own:
//   * a warp owns 64 columns (two adjacent columns per lane) and marches down the rows of its chunk;
//   * lane 0 feeds the warp's private ring (4 stages x 4 rows) with TMA 2-D boxes of 4 halo'd row
//     segments (hardware zero fill outside the domain), one mbarrier per stage;
//   * axis 2: 128-bit shared loads of the 2p+2 inputs of a column pair, Toeplitz-interior
//     coefficients from the kernel-parameter constant bank (the sum form shares the pair sums
//     x[i-k] + x[i+k] of its symmetric M and K rows); the few boundary columns of a domain read
//     their rows from global memory;
//   * axis 1: shift-form partial sums (poms_matvec3d_v3.cuh), 2 x (2p+1) accumulators per thread;
//   * no CTA barrier anywhere in the march; fused epilogues and the deterministic reduction as in 3-D.
#pragma once
#include <cuda.h>

struct MV2T {
    MV2 a;
    double t1m[11], t1k[11], t2m[11], t2k[11];   // interior (Toeplitz) band rows of axes 1, 2
    int lo1, hi1, lo2, hi2;                      // rows [lo, hi) of each axis equal to them
    int sym;                                     // sum form: interior rows symmetric (shared pair sums)
};

POMS_HIDDEN int poms_mv2_tma_launch(const CUtensorMap& tm, const MV2T& g, int p, int form, int epi, dim3 grid,
                                    cudaStream_t st);

#if POMS_TU == 7
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* tmap, int c0, int c1, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n" ::"r"(
            (unsigned)__cvta_generic_to_shared(smem_dst)),
        "l"(tmap), "r"(c0), "r"(c1), "r"((unsigned)__cvta_generic_to_shared(bar))
        : "memory");
}

template <int P>
struct MV2TCfg {
    static constexpr int W = 2 * P + 1, SH = P & 1;
    static constexpr int TW = 64;                          // columns per warp
    static constexpr int CW = TW + 2 * P + 2 * SH;         // box width (even)
    static constexpr int RB = 4, NSTG = 4, NWARP = 4;      // rows per box, ring depth, warps per CTA
    static constexpr int STG_D = RB * CW;                  // doubles per stage (bytes: multiple of 128)
    static constexpr size_t smem_bytes() { return (size_t)NWARP * NSTG * STG_D * 8 + NWARP * NSTG * 8 + 32 * 8; }
};

template <int P, int FORM, int EPI>
__global__ void __launch_bounds__(128, (P <= 3 ? 5 : 4))
kron_matvec2d_tma_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ MV2T g) {
    using C = MV2TCfg<P>;
    constexpr int W = C::W, SH = C::SH, TW = C::TW, CW = C::CW, RB = C::RB, NSTG = C::NSTG, STG_D = C::STG_D;
    constexpr bool TWO = (FORM == POMS_FORM_SUM);
    constexpr int NX = 2 * P + 2, E = 2;
    const MV2& a = g.a;
    extern __shared__ __align__(1024) unsigned char smem_raw2[];
    double* const ring_all = reinterpret_cast<double*>(smem_raw2);
    uint64_t* const mbar_all = reinterpret_cast<uint64_t*>(ring_all + (size_t)C::NWARP * NSTG * STG_D);
    double* const red = reinterpret_cast<double*>(mbar_all + C::NWARP * NSTG);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    double* const ring = ring_all + (size_t)wid * NSTG * STG_D;
    uint64_t* const mbar = mbar_all + wid * NSTG;

    const int i2_0 = (blockIdx.x * C::NWARP + wid) * TW - SH;      // first column of the warp's strip
    const int c_lo = blockIdx.y * a.chunk;
    const int c_hi = min(a.n1, c_lo + a.chunk);
    const bool warp_live = i2_0 < a.n2 && c_lo < a.n1;
    const int c0 = i2_0 + 2 * lane;                                 // this lane's two columns c0, c0+1
    const bool ok0 = c0 >= 0 && c0 < a.n2, ok1 = c0 + 1 >= 0 && c0 + 1 < a.n2;
    const bool toep2 = c0 >= g.lo2 && c0 + 2 <= g.hi2;
    const bool sym = g.sym != 0;

    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < NSTG; ++s) mbar_init(mbar + s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    }
    __syncwarp();

    double dsum = 0.0;
    if (warp_live) {
        double acc[E][W];
#pragma unroll
        for (int e = 0; e < E; ++e)
#pragma unroll
            for (int k = 0; k < W; ++k) acc[e][k] = 0.0;
        double dA0 = 1.0, dA1 = 1.0, dB0 = 0.0, dB1 = 0.0;        // diag factors of axis 2 (Jacobi epilogues)
        if (EPI == POMS_EPI_JACOBI || EPI == POMS_EPI_DINV) {
            dA0 = ok0 ? a.m2[(int64_t)c0 * W + P] : 1.0;
            dA1 = ok1 ? a.m2[(int64_t)(c0 + 1) * W + P] : 1.0;
            if (TWO) {
                dB0 = ok0 ? a.k2[(int64_t)c0 * W + P] : 0.0;
                dB1 = ok1 ? a.k2[(int64_t)(c0 + 1) * W + P] : 0.0;
            }
        }
        // input rows [jv0, jv1) exist; iterations j = jv0 .. jend-1 (the last ones may have no input)
        const int jv0 = max(c_lo - P, -a.glo), jv1 = min(c_hi + P, a.n1 + a.ghi), jend = c_hi + P;
        const int nbatch = (jend - jv0 + RB - 1) / RB;
        auto issue = [&](const int q) {       // lane 0: box of batch q into stage q % NSTG
            const int s = q % NSTG;
            mbar_expect_tx(mbar + s, RB * CW * 8);
            tma_load_2d(ring + (size_t)s * STG_D, &tmap, i2_0 - P, jv0 + q * RB + a.glo, mbar + s);
        };
        if (lane == 0) {
#pragma unroll
            for (int q = 0; q < NSTG; ++q)
                if (q < nbatch && jv0 + q * RB < jv1) issue(q);
        }
        unsigned phase = 0;
        double* yp = a.y + ((int64_t)c_lo * a.ld + c0);           // next output row, this lane's columns
        const double* bp = a.b ? a.b + ((int64_t)c_lo * a.ld + c0) : nullptr;
        const double* xq = a.x + ((int64_t)c_lo * a.ld + c0);
        constexpr bool NEED_B = (EPI != POMS_EPI_STORE);
        const bool need_b = NEED_B && a.b != nullptr;
        const bool need_x = (EPI == POMS_EPI_STORE && a.dot_out) || EPI == POMS_EPI_JACOBI;

#pragma unroll 1
        for (int q = 0; q < nbatch; ++q) {
            const int s = q % NSTG;
            const int j0 = jv0 + q * RB;
            const bool loaded = j0 < jv1;
            if (loaded) {
                mbar_wait(mbar + s, (phase >> s) & 1u);
                phase ^= (1u << s);
            }
            const double* const sbase = ring + (size_t)s * STG_D + 2 * lane;
#pragma unroll 1
            for (int r = 0; r < RB; ++r) {
                const int j = j0 + r;
                if (j >= jend) break;
                const bool have = j < jv1;
                const int i1 = j - P;
                const bool emit = i1 >= c_lo;
                // epilogue operands of the output row, requested before the arithmetic (a register
                // pipeline two rows deep was measured: SLOWER under the 128-register cap, C4 residual
                // 0.655 vs 0.551 ms; ncu still shows these loads exposed: long_scoreboard 3.0 per issue)
                double b0 = 0.0, b1 = 0.0, x0 = 0.0, x1 = 0.0;
                if (emit) {
                    if (need_b) {
                        if (ok0) b0 = bp[0];      // plain loads: b may alias y (in-place update)
                        if (ok1) b1 = bp[1];
                    }
                    if (need_x) {
                        if (ok0) x0 = __ldg(xq);
                        if (ok1) x1 = __ldg(xq + 1);
                    }
                }
                double ta[E] = {0.0, 0.0}, tb[E] = {0.0, 0.0}, vout[E];
                if (have) {
                    double xr[NX];
                    const double2* src = reinterpret_cast<const double2*>(sbase + r * CW);
#pragma unroll
                    for (int t = 0; t < NX / 2; ++t) {
                        const double2 v2 = src[t];
                        xr[2 * t] = v2.x;
                        xr[2 * t + 1] = v2.y;
                    }
                    if (toep2) {
                        if (TWO && sym) {
                            ta[0] = g.t2m[P] * xr[P];
                            ta[1] = g.t2m[P] * xr[P + 1];
                            tb[0] = g.t2k[P] * xr[P];
                            tb[1] = g.t2k[P] * xr[P + 1];
#pragma unroll
                            for (int k = 1; k <= P; ++k) {
                                const double sa = xr[P - k] + xr[P + k];
                                const double sb = xr[P + 1 - k] + xr[P + 1 + k];
                                ta[0] = fma(g.t2m[P + k], sa, ta[0]);
                                tb[0] = fma(g.t2k[P + k], sa, tb[0]);
                                ta[1] = fma(g.t2m[P + k], sb, ta[1]);
                                tb[1] = fma(g.t2k[P + k], sb, tb[1]);
                            }
                        } else {
#pragma unroll
                            for (int k = 0; k < W; ++k) {
                                ta[0] = fma(g.t2m[k], xr[k], ta[0]);
                                ta[1] = fma(g.t2m[k], xr[k + 1], ta[1]);
                                if (TWO) {
                                    tb[0] = fma(g.t2k[k], xr[k], tb[0]);
                                    tb[1] = fma(g.t2k[k], xr[k + 1], tb[1]);
                                }
                            }
                        }
                    } else {
                        // boundary columns of the domain: rows of M2 / K2 from global memory (L1 resident)
#pragma unroll
                        for (int k = 0; k < W; ++k) {
                            const double ma = ok0 ? __ldg(a.m2 + (int64_t)c0 * W + k) : 0.0;
                            const double mb = ok1 ? __ldg(a.m2 + (int64_t)(c0 + 1) * W + k) : 0.0;
                            ta[0] = fma(ma, xr[k], ta[0]);
                            ta[1] = fma(mb, xr[k + 1], ta[1]);
                            if (TWO) {
                                const double ka = ok0 ? __ldg(a.k2 + (int64_t)c0 * W + k) : 0.0;
                                const double kb = ok1 ? __ldg(a.k2 + (int64_t)(c0 + 1) * W + k) : 0.0;
                                tb[0] = fma(ka, xr[k], tb[0]);
                                tb[1] = fma(kb, xr[k + 1], tb[1]);
                            }
                        }
                    }
                }
                // axis 1: sliding partial sums.  FORM_SUM: y = K1 (M2 x) + M1 (K2 x)
                const bool toep1 = have && (j - P >= g.lo1) && (j + P < g.hi1);
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
                if (emit) {
                    double dg1 = 0.0, dg2 = 0.0;
                    if (EPI == POMS_EPI_JACOBI || EPI == POMS_EPI_DINV) {
                        dg1 = TWO ? __ldg(a.k1 + (int64_t)i1 * W + P) : __ldg(a.m1 + (int64_t)i1 * W + P);
                        dg2 = TWO ? __ldg(a.m1 + (int64_t)i1 * W + P) : 0.0;
                    }
#pragma unroll
                    for (int e = 0; e < E; ++e) {
                        const bool ok = e == 0 ? ok0 : ok1;
                        if (!ok) continue;
                        const double v = vout[e], bv = e == 0 ? b0 : b1, xv = e == 0 ? x0 : x1;
                        if (EPI == POMS_EPI_STORE) {
                            yp[e] = v;
                            if (need_x) dsum = fma(xv, v, dsum);
                        } else if (EPI == POMS_EPI_RESID) {
                            const double rr = bv - v;
                            yp[e] = rr;
                            dsum = fma(rr, rr, dsum);
                        } else if (EPI == POMS_EPI_AXPY) {
                            const double w_ = a.omega * v;
                            yp[e] = need_b ? bv + w_ : w_;
                            dsum = fma(w_, w_, dsum);
                        } else {
                            const double dA = e == 0 ? dA0 : dA1, dB = e == 0 ? dB0 : dB1;
                            const double dg = TWO ? dg1 * dA + dg2 * dB : dg1 * dA;
                            const double dr = a.omega * (bv - v) / dg;
                            yp[e] = (EPI == POMS_EPI_JACOBI) ? xv + dr : dr;
                            dsum = fma(dr, dr, dsum);
                        }
                    }
                    yp += a.ld;
                    if (bp) bp += a.ld;
                    xq += a.ld;
                }
            }
            // every lane is done with stage s: refill it with batch q + NSTG
            __syncwarp();
            if (lane == 0 && q + NSTG < nbatch && jv0 + (q + NSTG) * RB < jv1) issue(q + NSTG);
        }
    }
    if (a.dot_out) {
        const double tot = block_sum(dsum, red);
        const unsigned nbk = gridDim.x * gridDim.y;
        const unsigned bid = blockIdx.y * gridDim.x + blockIdx.x;
        grid_sum_finish(tot, a.dot_out, a.ws, nbk, bid, red);
    }
}

template <int P, int FORM, int EPI>
static int launch_mv2_tma_inst(const CUtensorMap& tm, const MV2T& g, dim3 grid, cudaStream_t st) {
    const size_t smem = MV2TCfg<P>::smem_bytes();
    auto kern = kron_matvec2d_tma_kernel<P, FORM, EPI>;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return fail_cuda(e, "cudaFuncSetAttribute(mv2 tma)");
        attr_set = true;
    }
    kern<<<grid, 128, smem, st>>>(tm, g);
    return 0;
}
template <int P, int FORM>
static int launch_mv2_tma_epi(const CUtensorMap& tm, const MV2T& g, int epi, dim3 grid, cudaStream_t st) {
    switch (epi) {
        case POMS_EPI_STORE: return launch_mv2_tma_inst<P, FORM, POMS_EPI_STORE>(tm, g, grid, st);
        case POMS_EPI_RESID: return launch_mv2_tma_inst<P, FORM, POMS_EPI_RESID>(tm, g, grid, st);
        case POMS_EPI_JACOBI: return launch_mv2_tma_inst<P, FORM, POMS_EPI_JACOBI>(tm, g, grid, st);
        case POMS_EPI_DINV: return launch_mv2_tma_inst<P, FORM, POMS_EPI_DINV>(tm, g, grid, st);
        case POMS_EPI_AXPY: return launch_mv2_tma_inst<P, FORM, POMS_EPI_AXPY>(tm, g, grid, st);
        default: return bad_arg(15, "epilogue");
    }
}
template <int P>
static int launch_mv2_tma_p(const CUtensorMap& tm, const MV2T& g, int form, int epi, dim3 grid, cudaStream_t st) {
    if (form == POMS_FORM_SINGLE) return launch_mv2_tma_epi<P, POMS_FORM_SINGLE>(tm, g, epi, grid, st);
    return launch_mv2_tma_epi<P, POMS_FORM_SUM>(tm, g, epi, grid, st);
}
int poms_mv2_tma_launch(const CUtensorMap& tm, const MV2T& g, int p, int form, int epi, dim3 grid, cudaStream_t st) {
    switch (p) {
        case 1: return launch_mv2_tma_p<1>(tm, g, form, epi, grid, st);
        case 2: return launch_mv2_tma_p<2>(tm, g, form, epi, grid, st);
        case 3: return launch_mv2_tma_p<3>(tm, g, form, epi, grid, st);
        case 4: return launch_mv2_tma_p<4>(tm, g, form, epi, grid, st);
        case 5: return launch_mv2_tma_p<5>(tm, g, form, epi, grid, st);
        default: return bad_arg(9, "p must be 1..5");
    }
}
#endif  // POMS_TU == 7

#if POMS_TU == 0
static int g_mv2_variant = -1;     // 1 = warp-autonomous TMA kernel (default), 0 = round-1 kernel
extern "C" void poms_set_matvec2d_variant(int v) { g_mv2_variant = v; }
static int mv2_variant() {
    if (g_mv2_variant < 0) {
        const char* e = getenv("POMS_B200_MV2_VARIANT");
        g_mv2_variant = e ? atoi(e) : 1;
    }
    return g_mv2_variant;
}

// returns 0 on success, 1 if the TMA path does not apply (caller runs the round-1 kernel)
static int try_matvec2d_tma(const MV2& a0, int p, int form, int epilogue, const double* toep, const int* toep_rng,
                            cudaStream_t st) {
    if (mv2_variant() == 0) return 1;
    if (((uintptr_t)a0.x & 15) || (a0.ld & 1)) return 1;
    if (a0.n2 < 8) return 1;
    PFN_encodeTiled enc = get_encode_tiled();
    if (!enc) return 1;
    const int W = 2 * p + 1, sh = p & 1;
    MV2T g;
    g.a = a0;
    for (int k = 0; k < 11; ++k) g.t1m[k] = g.t1k[k] = g.t2m[k] = g.t2k[k] = 0.0;
    g.lo1 = g.hi1 = g.lo2 = g.hi2 = 0;
    g.sym = 0;
    // 2-D tensor map over the rows incl. ghost rows; box = 4 halo'd row segments of one warp
    CUtensorMap tm;
    {
        static std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> cache2;
        static std::mutex mu2;
        const int boxw = 64 + 2 * p + 2 * sh;
        const TmapKey key{(uint64_t)(uintptr_t)(a0.x - (int64_t)a0.glo * a0.ld), (uint64_t)a0.n2, 0,
                          (uint64_t)(a0.n1 + a0.glo + a0.ghi), (uint64_t)a0.ld, 0, ((uint64_t)boxw << 32) | 4u};
        std::lock_guard<std::mutex> lk(mu2);
        auto it = cache2.find(key);
        if (it != cache2.end()) {
            tm = it->second;
        } else {
            cuuint64_t dims[2] = {(cuuint64_t)a0.n2, (cuuint64_t)(a0.n1 + a0.glo + a0.ghi)};
            cuuint64_t strides[1] = {(cuuint64_t)a0.ld * 8};
            cuuint32_t box[2] = {(cuuint32_t)boxw, 4};
            cuuint32_t es[2] = {1, 1};
            CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void*)(a0.x - (int64_t)a0.glo * a0.ld), dims,
                             strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                             CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) return 1;
            if (cache2.size() > 4096) cache2.clear();
            cache2.emplace(key, tm);
        }
    }
    const int nwarp = (a0.n2 + sh + 63) / 64;
    const int g2 = (nwarp + 3) / 4;
    g.a.chunk = pick_chunk2d(a0.n1, g2, p);
    const int g1 = (a0.n1 + g.a.chunk - 1) / g.a.chunk;
    if ((int64_t)g2 * g1 > POMS_MAX_PARTIALS) return 1;
    // Toeplitz-interior rows of the bands come from the caller (host hints); without them every
    // column / row takes the global-memory coefficient path.  toep: [axis(0,1)][m|k][W], rng {lo1,hi1,lo2,hi2}
    if (toep && toep_rng) {
        for (int k = 0; k < W; ++k) {
            g.t1m[k] = toep[(0 * 2 + 0) * W + k];
            g.t1k[k] = toep[(0 * 2 + 1) * W + k];
            g.t2m[k] = toep[(1 * 2 + 0) * W + k];
            g.t2k[k] = toep[(1 * 2 + 1) * W + k];
        }
        g.lo1 = toep_rng[0];
        g.hi1 = toep_rng[1];
        g.lo2 = toep_rng[2];
        g.hi2 = toep_rng[3];
        bool sym = form == POMS_FORM_SUM;
        double* rows[2] = {g.t2m, g.t2k};
        for (int r = 0; r < 2 && sym; ++r) {
            double mx = 0.0;
            for (int k = 0; k < W; ++k) mx = fmax(mx, fabs(rows[r][k]));
            for (int k = 0; k < W; ++k) sym = sym && fabs(rows[r][k] - rows[r][W - 1 - k]) <= 1e-13 * mx;
        }
        g.sym = sym ? 1 : 0;
        // assembled rows are symmetric to rounding only: the pair-sum form uses the MEAN of the two
        // halves, so that row sums (K 1 = 0, partition of unity) carry no systematic bias
        if (sym)
            for (int r = 0; r < 2; ++r)
                for (int k = 1; k <= p; ++k) rows[r][p + k] = rows[r][p - k] = 0.5 * (rows[r][p + k] + rows[r][p - k]);
    }
    return poms_mv2_tma_launch(tm, g, p, form, epilogue, dim3(g2, g1), st);
}
#endif
