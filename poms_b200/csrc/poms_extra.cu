// poms_extra.cu -- self-contained units of libpoms_b200.so (own kernels + C ABI, see
// include/poms_b200.h):
//   * peer-memory halo exchange over NVLink (CUDA IPC): replaces the NCCL send/recv pair behind
//     `update_ghost_regions` (/root/reference/sources/kron_product.py:76,87, solvers.py:162,215)
//   * dense per-axis contraction on the fp64 tensor cores (DMMA)
//   * full (non-separable) 3-D stencil mat-vec, two-colour Jacobi update
#include <string.h>
// shared helpers (error text, launch counter, deterministic grid reduction): the common part of
// poms_kernels.cu, none of its translation-unit sections
#define POMS_TU 99
#include "poms_kernels.cu"

static int x_fail_cuda(cudaError_t e, const char* where) {
    snprintf(g_err, sizeof(g_err), "%s: %s", where, cudaGetErrorString(e));
    return (int)e;
}
static int x_bad_arg(int idx, const char* what) {
    snprintf(g_err, sizeof(g_err), "bad argument %d: %s", idx, what);
    return -idx;
}

// ------------------------------------------------------------------------------------------
// CUDA IPC plumbing: one process per GPU; a rank maps the arenas of its two slab neighbours
// ------------------------------------------------------------------------------------------
extern "C" int poms_ipc_alloc(int64_t bytes, void** ptr_out) {
    if (bytes <= 0) return x_bad_arg(1, "bytes");
    if (!ptr_out) return x_bad_arg(2, "ptr_out");
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, (size_t)bytes);
    if (e != cudaSuccess) return x_fail_cuda(e, "poms_ipc_alloc(cudaMalloc)");
    e = cudaMemset(p, 0, (size_t)bytes);
    if (e != cudaSuccess) return x_fail_cuda(e, "poms_ipc_alloc(cudaMemset)");
    *ptr_out = p;
    return 0;
}
extern "C" int poms_ipc_free(void* ptr) {
    cudaError_t e = cudaFree(ptr);
    return e == cudaSuccess ? 0 : x_fail_cuda(e, "poms_ipc_free");
}
extern "C" int poms_ipc_handle_bytes(void) { return (int)sizeof(cudaIpcMemHandle_t); }
extern "C" int poms_ipc_get_handle(void* ptr, void* handle_out) {
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, ptr);
    if (e != cudaSuccess) return x_fail_cuda(e, "poms_ipc_get_handle");
    memcpy(handle_out, &h, sizeof(h));
    return 0;
}
extern "C" int poms_ipc_open(const void* handle, void** ptr_out) {
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    void* p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) return x_fail_cuda(e, "poms_ipc_open");
    *ptr_out = p;
    return 0;
}
extern "C" int poms_ipc_close(void* ptr) {
    cudaError_t e = cudaIpcCloseMemHandle(ptr);
    return e == cudaSuccess ? 0 : x_fail_cuda(e, "poms_ipc_close");
}

// ------------------------------------------------------------------------------------------
// Halo exchange by peer stores.  Every rank pushes its outermost `width` owned planes straight
// into the ghost planes of its neighbours (NVLink peer writes), so there is no staging buffer,
// no pack kernel and no NCCL proxy: ONE small kernel per exchange.
//
// flags (uint64, in the rank's own IPC-shared memory; neighbours write into them):
//   [0] sequence number of the last completed exchange (own; lives on the device so that the
//       kernel can be replayed from a CUDA graph)
//   [1] / [2]  "entered exchange s" written by the lower / upper neighbour: everything queued
//       before its exchange kernel has finished, so its ghost planes may be overwritten
//   [3] / [4]  "data of exchange s has landed" written by the lower / upper neighbour
//   [5] block ticket
// Protocol of exchange s on every rank: signal ENTER to both neighbours; wait for their ENTER;
// push; fence; the last block signals DATA and waits for the neighbours' DATA before it exits,
// so the next kernel in the stream sees complete ghost planes.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void st_release_sys(uint64_t* p, uint64_t v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;\n" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ uint64_t ld_acquire_sys(const uint64_t* p) {
    uint64_t v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];\n" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void spin_until(const uint64_t* p, uint64_t seq) {
    while (ld_acquire_sys(p) < seq) __nanosleep(20);
}

__global__ void __launch_bounds__(256) halo_push_kernel(
    const double* __restrict__ src_lo, double* __restrict__ dst_lo,   // my lowest owned planes -> lower neighbour
    const double* __restrict__ src_hi, double* __restrict__ dst_hi,   // my highest owned planes -> upper neighbour
    int64_t n2, /* double2 elements per direction */
    uint64_t* my_flags, uint64_t* lo_flags, uint64_t* hi_flags) {
    __shared__ uint64_t s_seq;
    if (threadIdx.x == 0) {
        const uint64_t seq = ld_acquire_sys(my_flags + 0) + 1;
        if (blockIdx.x == 0) {
            if (lo_flags) st_release_sys(lo_flags + 2, seq);   // I am the lower rank's UPPER neighbour
            if (hi_flags) st_release_sys(hi_flags + 1, seq);
        }
        if (lo_flags) spin_until(my_flags + 1, seq);
        if (hi_flags) spin_until(my_flags + 2, seq);
        s_seq = seq;
    }
    __syncthreads();
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t t0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    // four 16-byte loads in flight per thread before the peer stores (NVLink writes are posted)
    auto push = [&](const double* src, double* dst) {
        const double2* s = reinterpret_cast<const double2*>(src);
        double2* d = reinterpret_cast<double2*>(dst);
        int64_t i = t0;
        for (; i + 3 * stride < n2; i += 4 * stride) {
            const double2 a0 = s[i], a1 = s[i + stride], a2 = s[i + 2 * stride], a3 = s[i + 3 * stride];
            d[i] = a0;
            d[i + stride] = a1;
            d[i + 2 * stride] = a2;
            d[i + 3 * stride] = a3;
        }
        for (; i < n2; i += stride) d[i] = s[i];
    };
    if (lo_flags) push(src_lo, dst_lo);
    if (hi_flags) push(src_hi, dst_hi);
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint64_t seq = s_seq;
        unsigned* ticket = reinterpret_cast<unsigned*>(my_flags + 5);
        const unsigned t = atomicAdd(ticket, 1u);
        if (t == gridDim.x - 1) {
            __threadfence_system();
            *ticket = 0u;
            if (lo_flags) st_release_sys(lo_flags + 4, seq);
            if (hi_flags) st_release_sys(hi_flags + 3, seq);
            if (lo_flags) spin_until(my_flags + 3, seq);
            if (hi_flags) spin_until(my_flags + 4, seq);
            st_release_sys(my_flags + 0, seq);
        }
    }
}

extern "C" int poms_halo_flags_bytes(void) { return 64; }

extern "C" int poms_halo_exchange_p2p(const double* src_lo, double* dst_lo, const double* src_hi, double* dst_hi,
                                      int64_t n_doubles, void* my_flags, void* lo_flags, void* hi_flags,
                                      void* stream) {
    if (!my_flags) return x_bad_arg(6, "my_flags");
    if (lo_flags && (!src_lo || !dst_lo)) return x_bad_arg(1, "lower neighbour pointers");
    if (hi_flags && (!src_hi || !dst_hi)) return x_bad_arg(3, "upper neighbour pointers");
    if (n_doubles < 0 || (n_doubles & 1)) return x_bad_arg(5, "n_doubles must be even (16-byte copies)");
    if ((((uintptr_t)src_lo | (uintptr_t)dst_lo | (uintptr_t)src_hi | (uintptr_t)dst_hi) & 15) != 0)
        return x_bad_arg(1, "halo blocks must be 16-byte aligned");
    if (!lo_flags && !hi_flags) return 0;
    const int64_t n2 = n_doubles / 2;
    int blocks = (int)((n2 + 256 * 8 - 1) / (256 * 8));   // ~8 x 16 B per thread
    if (blocks < 1) blocks = 1;
    if (blocks > 132) blocks = 132;                        // all blocks resident at once on 148 SMs
    halo_push_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(src_lo, dst_lo, src_hi, dst_hi, n2,
                                                              (uint64_t*)my_flags, (uint64_t*)lo_flags,
                                                              (uint64_t*)hi_flags);
    g_launches++;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return x_fail_cuda(e, "poms_halo_exchange_p2p");
    return 0;
}

// ------------------------------------------------------------------------------------------
// Dense per-axis contraction on the fp64 tensor cores (DMMA, mma.sync m8n8k4):
//     out[o, i, c] = sum_j Q[i, j] * in[o, j, c]        (Q dense n_out x n_in, row-major)
// the 1-D eigenbasis contractions of the fast-diagonalisation solve that replaces
// splu(csc_matrix(Ac)).solve(rc) (/root/reference/sources/mg_jac.py:98-99) and the dense
// Kronecker solve of /root/reference/sources/kron_product.py:93-117.  A/B partner of
// poms_axis_gather with W = n_in (scalar FMA, one coefficient load per FMA).
// CTA = 4 warps, tile 64 (i) x 64 (c) with K chunks of 32 staged in shared memory; a warp owns
// 16 rows x 64 columns = 16 accumulator fragments.  TRANS: the contraction runs along the
// CONTIGUOUS axis (n_inner == 1): the "column" index is then the outer index o.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void dmma8x8x4(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};\n"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}

template <bool TRANS>
__global__ void __launch_bounds__(128) axis_dense_dmma_kernel(
    const double* __restrict__ in, double* __restrict__ out, const double* __restrict__ Q, int n_in, int n_out,
    int64_t so_in, int64_t sa_in, int64_t so_out, int64_t sa_out, int64_t n_cols) {
    // CTA tile 64 (i) x 64 (c); a warp owns 16 rows x 64 columns = 2 x 8 accumulator fragments, so every
    // B fragment feeds two DMMAs (10 shared loads per 16 DMMAs; the 32-row version needed 9 per 8)
    constexpr int TM = 64, TN = 64, TK = 32, QP = TK + 4, XP = TN + 8;
    __shared__ double Qs[TM][QP];     // Q[i0 + r][k0 + k]
    __shared__ double Xs[TK][XP];     // X[k0 + k][c0 + c]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int i0 = blockIdx.y * TM;
    const int64_t c0 = (int64_t)blockIdx.x * TN;
    const int64_t o = TRANS ? 0 : blockIdx.z;
    const double* inb = in + o * so_in;
    double* outb = out + o * so_out;
    double acc[2][8][2];
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int t = 0; t < 8; ++t) acc[h][t][0] = acc[h][t][1] = 0.0;
    for (int k0 = 0; k0 < n_in; k0 += TK) {
        for (int e = tid; e < TM * TK; e += 128) {
            const int r = e / TK, k = e - r * TK;
            const int i = i0 + r, j = k0 + k;
            Qs[r][k] = (i < n_out && j < n_in) ? __ldg(Q + (int64_t)i * n_in + j) : 0.0;
        }
        if (!TRANS) {
            for (int e = tid; e < TK * TN; e += 128) {
                const int k = e / TN, c = e - k * TN;
                const int j = k0 + k;
                const int64_t cc = c0 + c;
                Xs[k][c] = (j < n_in && cc < n_cols) ? inb[(int64_t)j * sa_in + cc] : 0.0;
            }
        } else {
            for (int e = tid; e < TK * TN; e += 128) {
                const int c = e / TK, k = e - c * TK;      // k fastest: contiguous in memory
                const int j = k0 + k;
                const int64_t cc = c0 + c;
                Xs[k][c] = (j < n_in && cc < n_cols) ? in[cc * so_in + j] : 0.0;
            }
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < TK; kk += 4) {
            const double a0 = Qs[warp * 16 + (lane >> 2)][kk + (lane & 3)];
            const double a1 = Qs[warp * 16 + 8 + (lane >> 2)][kk + (lane & 3)];
#pragma unroll
            for (int t = 0; t < 8; ++t) {
                const double b = Xs[kk + (lane & 3)][t * 8 + (lane >> 2)];
                dmma8x8x4(acc[0][t][0], acc[0][t][1], a0, b);
                dmma8x8x4(acc[1][t][0], acc[1][t][1], a1, b);
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int h2 = 0; h2 < 2; ++h2) {
        const int i = i0 + warp * 16 + 8 * h2 + (lane >> 2);
        if (i < n_out) {
#pragma unroll
            for (int t = 0; t < 8; ++t) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int64_t cc = c0 + t * 8 + 2 * (lane & 3) + h;
                    if (cc < n_cols) {
                        if (!TRANS) outb[(int64_t)i * sa_out + cc] = acc[h2][t][h];
                        else out[cc * so_out + i] = acc[h2][t][h];
                    }
                }
            }
        }
    }
}

extern "C" int poms_axis_dense_dmma(const double* in, double* out, const double* Q, int n_in, int n_out,
                                    int64_t n_outer, int64_t so_in, int64_t sa_in, int64_t so_out,
                                    int64_t sa_out, int64_t n_inner, void* stream) {
    if (!in || !out || !Q) return x_bad_arg(1, "null pointer");
    if (n_in < 1 || n_out < 1 || n_outer < 1 || n_inner < 1) return x_bad_arg(4, "extent");
    cudaStream_t st = (cudaStream_t)stream;
    if (n_inner == 1) {
        // contraction along the contiguous axis: columns = the n_outer lines
        if (sa_in != 1 || sa_out != 1) return x_bad_arg(8, "contiguous axis needs unit stride");
        dim3 grid((unsigned)((n_outer + 63) / 64), (unsigned)((n_out + 63) / 64), 1);
        axis_dense_dmma_kernel<true><<<grid, 128, 0, st>>>(in, out, Q, n_in, n_out, so_in, 1, so_out, 1, n_outer);
    } else {
        if (n_outer > 65535) return x_bad_arg(6, "n_outer");
        dim3 grid((unsigned)((n_inner + 63) / 64), (unsigned)((n_out + 63) / 64), (unsigned)n_outer);
        axis_dense_dmma_kernel<false><<<grid, 128, 0, st>>>(in, out, Q, n_in, n_out, so_in, sa_in, so_out,
                                                           sa_out, n_inner);
    }
    g_launches++;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return x_fail_cuda(e, "poms_axis_dense_dmma");
    return 0;
}

// ------------------------------------------------------------------------------------------
// Full (non-separable) 3-D stencil mat-vec: y[i] = sum_k S[i, k1, k2, k3] x[i + k - p]
// = spl StencilMatrix.dot in 3-D (slides/content.tex:285-290; the operator the reference's solvers
// call, sources/solvers.py:85,103,209).  S is (n1, n2, n3, 2p1+1, 2p2+1, 2p3+1) row-major: (2p+1)^3
// coefficients PER ROW (343 doubles = 2.7 KB at p = 3), so the kernel is bound by the coefficient
// stream, not by x: one WARP per output point, lanes stride over the point's contiguous coefficient
// block (coalesced 256-byte requests), x comes from L1/L2, warp-shuffle sum.
// ------------------------------------------------------------------------------------------
template <int EPI>
__global__ void __launch_bounds__(256) stencil_matvec3d_kernel(
    const double* __restrict__ x, double* __restrict__ y, const double* __restrict__ b,
    const double* __restrict__ S, int n1, int n2, int n3, int64_t ld, int64_t pld, int glo, int ghi, int p1, int p2,
    int p3, double omega, double* dot_out, void* ws) {
    __shared__ double red[32];
    extern __shared__ int soff[];          // per coefficient: offset of its x entry relative to the point
    const int W2 = 2 * p2 + 1, W3 = 2 * p3 + 1, NC = (2 * p1 + 1) * W2 * W3;
    const int lane = threadIdx.x & 31;
    for (int c = threadIdx.x; c < NC; c += blockDim.x) {
        const int k3 = c % W3, k2 = (c / W3) % W2, k1 = c / (W3 * W2);
        soff[c] = (int)((k1 - p1) * pld + (k2 - p2) * ld + (k3 - p3));
    }
    __syncthreads();
    const int64_t npts = (int64_t)n1 * n2 * n3;
    const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    double dsum = 0.0;
    for (int64_t pt = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); pt < npts; pt += nwarps) {
        const int i3 = (int)(pt % n3);
        const int i2 = (int)((pt / n3) % n2);
        const int i1 = (int)(pt / ((int64_t)n3 * n2));
        const double* Sp = S + pt * NC;
        const double* xp = x + ((int64_t)i1 * pld + (int64_t)i2 * ld + i3);
        double v = 0.0;
        // interior points (the whole stencil inside the stored planes): no bounds checks
        const bool inner = i1 - p1 >= -glo && i1 + p1 < n1 + ghi && i2 >= p2 && i2 + p2 < n2 && i3 >= p3 &&
                           i3 + p3 < n3;
        if (inner) {
            double v1 = 0.0;
            int c = lane;
            for (; c + 32 < NC; c += 64) {
                v = fma(__ldg(Sp + c), xp[soff[c]], v);
                v1 = fma(__ldg(Sp + c + 32), xp[soff[c + 32]], v1);
            }
            if (c < NC) v = fma(__ldg(Sp + c), xp[soff[c]], v);
            v += v1;
        } else {
            for (int c = lane; c < NC; c += 32) {
                const int k3 = c % W3, k2 = (c / W3) % W2, k1 = c / (W3 * W2);
                const int j1 = i1 + k1 - p1, j2 = i2 + k2 - p2, j3 = i3 + k3 - p3;
                if (j1 >= -glo && j1 < n1 + ghi && j2 >= 0 && j2 < n2 && j3 >= 0 && j3 < n3)
                    v = fma(__ldg(Sp + c), xp[soff[c]], v);
            }
        }
        v = warp_sum(v);
        if (lane == 0) {
            const int64_t o = (int64_t)i1 * pld + (int64_t)i2 * ld + i3;
            if (EPI == POMS_EPI_STORE) {
                y[o] = v;
                if (dot_out) dsum = fma(x[o], v, dsum);
            } else if (EPI == POMS_EPI_RESID) {
                const double rr = b[o] - v;
                y[o] = rr;
                dsum = fma(rr, rr, dsum);
            } else if (EPI == POMS_EPI_AXPY) {
                const double w_ = omega * v;
                y[o] = b ? b[o] + w_ : w_;
                dsum = fma(w_, w_, dsum);
            } else {
                const double dg = Sp[(p1 * W2 + p2) * W3 + p3];
                const double dr = omega * (b[o] - v) / dg;
                y[o] = (EPI == POMS_EPI_JACOBI) ? x[o] + dr : dr;
                dsum = fma(dr, dr, dsum);
            }
        }
    }
    if (dot_out) {
        const double tot = block_sum(dsum, red);
        grid_sum_finish(tot, dot_out, ws, gridDim.x, blockIdx.x, red);
    }
}

extern "C" int poms_stencil_matvec_3d(const double* x, double* y, const double* b, const double* S, int n1, int n2,
                                      int n3, int64_t ld, int64_t pld, int glo, int ghi, int p1, int p2, int p3,
                                      int epilogue, double omega, double* dot_out, void* ws, void* stream) {
    if (!x) return bad_arg(1, "x");
    if (!y) return bad_arg(2, "y");
    if (epilogue != POMS_EPI_STORE && epilogue != POMS_EPI_AXPY && !b) return bad_arg(3, "b required by epilogue");
    if (!S) return bad_arg(4, "S");
    if (n1 < 1 || n2 < 1 || n3 < 1) return bad_arg(5, "extent");
    if (ld < n3) return bad_arg(8, "ld");
    if (pld < ld * n2) return bad_arg(9, "pld");
    if (p1 < 0 || p2 < 0 || p3 < 0 || p1 > 5 || p2 > 5 || p3 > 5) return bad_arg(12, "pads");
    if (dot_out && !ws) return bad_arg(18, "ws");
    const int64_t npts = (int64_t)n1 * n2 * n3;
    int64_t blocks = (npts + 7) / 8;
    if (blocks > 148 * 16) blocks = 148 * 16;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t smem = sizeof(int) * (size_t)(2 * p1 + 1) * (2 * p2 + 1) * (2 * p3 + 1);
    if (pld * (int64_t)(p1 + 1) > 0x7fffffff) return bad_arg(9, "plane pitch too large for 32-bit stencil offsets");
#define POMS_S3(E) stencil_matvec3d_kernel<E><<<(int)blocks, 256, smem, st>>>(x, y, b, S, n1, n2, n3, ld, pld, glo, ghi, p1, p2, p3, omega, dot_out, ws)
    switch (epilogue) {
        case POMS_EPI_STORE: POMS_S3(POMS_EPI_STORE); break;
        case POMS_EPI_RESID: POMS_S3(POMS_EPI_RESID); break;
        case POMS_EPI_JACOBI: POMS_S3(POMS_EPI_JACOBI); break;
        case POMS_EPI_DINV: POMS_S3(POMS_EPI_DINV); break;
        case POMS_EPI_AXPY: POMS_S3(POMS_EPI_AXPY); break;
        default: return bad_arg(15, "epilogue");
    }
#undef POMS_S3
    CHECK_LAUNCH("poms_stencil_matvec_3d");
    return 0;
}

// ------------------------------------------------------------------------------------------
// Two-colour (red-black) update x[i] += d[i] on the points with (i1 + i2 [+ i3] + off) % 2 == colour:
// the half sweep of a red-black damped Jacobi smoother (named as future work in the reference's
// slides, slides/content.tex:393).  d = omega D^-1 (b - A x) comes from the fused DINV epilogue.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) color_add_kernel(double* __restrict__ x, const double* __restrict__ d, int n1,
                                                        int n2, int n3, int64_t ld, int64_t pld, int off, int colour) {
    const int64_t total = (int64_t)n1 * n2 * n3;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
        const int i3 = (int)(t % n3);
        const int i2 = (int)((t / n3) % n2);
        const int i1 = (int)(t / ((int64_t)n3 * n2));
        if (((i1 + i2 + i3 + off) & 1) == colour) {
            const int64_t o = (int64_t)i1 * pld + (int64_t)i2 * ld + i3;
            x[o] += d[o];
        }
    }
}
extern "C" int poms_color_add(double* x, const double* d, int n1, int n2, int n3, int64_t ld, int64_t pld, int off,
                              int colour, void* stream) {
    if (!x || !d) return bad_arg(1, "null pointer");
    if (n1 < 1 || n2 < 1 || n3 < 1) return bad_arg(3, "extent");
    const int64_t total = (int64_t)n1 * n2 * n3;
    int64_t blocks = (total + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    color_add_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(x, d, n1, n2, n3, ld, pld, off, colour & 1);
    CHECK_LAUNCH("poms_color_add");
    return 0;
}
