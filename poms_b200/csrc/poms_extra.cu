// poms_extra.cu -- self-contained units of libpoms_b200.so (own kernels + C ABI, see
// include/poms_b200.h):
//   * peer-memory halo exchange over NVLink (CUDA IPC): replaces the NCCL send/recv pair behind
//     `update_ghost_regions` (/root/reference/sources/kron_product.py:76,87, solvers.py:162,215)
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include "poms_b200.h"

#define POMS_HIDDEN __attribute__((visibility("hidden")))
extern POMS_HIDDEN thread_local char g_err[256];
extern POMS_HIDDEN int64_t g_launches;

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
// CTA = 4 warps, tile 32 (i) x 64 (c) with K chunks of 32 staged in shared memory; a warp owns
// 8 rows x 64 columns = 8 accumulator fragments.  TRANS: the contraction runs along the
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
    constexpr int TM = 32, TN = 64, TK = 32, QP = TK + 4, XP = TN + 8;
    __shared__ double Qs[TM][QP];     // Q[i0 + r][k0 + k]
    __shared__ double Xs[TK][XP];     // X[k0 + k][c0 + c]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int i0 = blockIdx.y * TM;
    const int64_t c0 = (int64_t)blockIdx.x * TN;
    const int64_t o = TRANS ? 0 : blockIdx.z;
    const double* inb = in + o * so_in;
    double* outb = out + o * so_out;
    double acc[8][2];
#pragma unroll
    for (int t = 0; t < 8; ++t) acc[t][0] = acc[t][1] = 0.0;
    for (int k0 = 0; k0 < n_in; k0 += TK) {
        // stage Q tile (coalesced along k) and X tile
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
            const double a = Qs[warp * 8 + (lane >> 2)][kk + (lane & 3)];
#pragma unroll
            for (int t = 0; t < 8; ++t) {
                const double b = Xs[kk + (lane & 3)][t * 8 + (lane >> 2)];
                dmma8x8x4(acc[t][0], acc[t][1], a, b);
            }
        }
        __syncthreads();
    }
    const int i = i0 + warp * 8 + (lane >> 2);
    if (i < n_out) {
#pragma unroll
        for (int t = 0; t < 8; ++t) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int64_t cc = c0 + t * 8 + 2 * (lane & 3) + h;
                if (cc < n_cols) {
                    if (!TRANS) outb[(int64_t)i * sa_out + cc] = acc[t][h];
                    else out[cc * so_out + i] = acc[t][h];
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
        dim3 grid((unsigned)((n_outer + 63) / 64), (unsigned)((n_out + 31) / 32), 1);
        axis_dense_dmma_kernel<true><<<grid, 128, 0, st>>>(in, out, Q, n_in, n_out, so_in, 1, so_out, 1, n_outer);
    } else {
        if (n_outer > 65535) return x_bad_arg(6, "n_outer");
        dim3 grid((unsigned)((n_inner + 63) / 64), (unsigned)((n_out + 31) / 32), (unsigned)n_outer);
        axis_dense_dmma_kernel<false><<<grid, 128, 0, st>>>(in, out, Q, n_in, n_out, so_in, sa_in, so_out,
                                                           sa_out, n_inner);
    }
    g_launches++;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return x_fail_cuda(e, "poms_axis_dense_dmma");
    return 0;
}
