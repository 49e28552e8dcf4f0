// cuda_emu.h -- TEST INFRASTRUCTURE ONLY: just enough of the CUDA execution model to compile a kernel
// source file for the HOST (tests/test_kernel_host_emulation.py): one OS thread per CUDA thread of a
// block, blocks one after the other, __syncthreads() = pthread barrier, __shared__ = function-local
// static.  Built with -fsanitize=address or -fsanitize=thread, so that out-of-bounds accesses and
// missing barriers of the kernel show up on a machine without a GPU.  Never part of libpoms_b200.so.
// Limits: (1) the threads of a warp run independently, so code that relies on implicit warp lockstep
// without __syncwarp() is reported as a race (conservative); (2) static shared arrays keep their content
// from one block to the next (an uninitialised read could be masked); dynamic shared memory is one
// global buffer whose tail beyond the requested size is poisoned under the address sanitizer, so a
// carve-up that overruns the launch's byte count is caught; (3) cp.async waits and the memory model
// below block scope are not modelled; (4) timing means nothing.
#pragma once
#include <pthread.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <thread>
#include <vector>

struct dim3 {
    unsigned x, y, z;
    dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
struct emu_idx {
    unsigned x, y, z;
};
static thread_local emu_idx threadIdx, blockIdx;
static dim3 blockDim, gridDim;
static pthread_barrier_t emu_bar;
static pthread_barrier_t emu_warp_bar[32];      // __syncwarp(): one barrier per warp of the block
static inline void __syncthreads() { pthread_barrier_wait(&emu_bar); }
static inline void __syncwarp() { pthread_barrier_wait(&emu_warp_bar[threadIdx.x >> 5]); }
typedef void* cudaStream_t;
typedef int cudaError_t;
constexpr int cudaSuccess = 0, cudaFuncAttributeMaxDynamicSharedMemorySize = 8;
template <class K>
static inline int cudaFuncSetAttribute(K, int, int) { return 0; }
constexpr int EMU_DYN_SMEM_DOUBLES = 28 * 1024;   // 224 KB, the most a kernel can ask for
alignas(1024) static unsigned char emu_dyn_smem[EMU_DYN_SMEM_DOUBLES * 8];   // `extern __shared__` of every kernel
#if defined(__SANITIZE_ADDRESS__)
extern "C" void __asan_poison_memory_region(void const volatile*, size_t);
extern "C" void __asan_unpoison_memory_region(void const volatile*, size_t);
#endif
static char g_err[256];
static long long g_launches;

#define __global__
#define __shared__ static
#define __launch_bounds__(...)
#define __host__
#define __device__
#define __forceinline__ inline
template <class T>
static inline T __ldg(const T* p) { return *p; }
template <class T>
static inline T __ldcs(const T* p) { return *p; }
struct alignas(16) double2 {
    double x, y;
};
static inline double2 make_double2(double x, double y) { return double2{x, y}; }
// warp shuffle: every lane of the warp deposits its value, warp barrier, reads its partner's
static double emu_shfl_buf[32][32];
static inline double __shfl_xor_sync(unsigned, double v, int o) {
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    emu_shfl_buf[w][lane] = v;
    __syncwarp();
    const double r = emu_shfl_buf[w][lane ^ o];
    __syncwarp();
    return r;
}
// mma.sync.aligned.m8n8k4.row.col.f64: D (8 x 8) += A (8 x 4, row major) * B (4 x 8, column major).  Fragment
// layout of the PTX ISA: lane l holds A[l / 4][l % 4], B[l % 4][l / 4] and D[l / 4][2 * (l % 4) + {0, 1}].
static double emu_mma_a[32][32], emu_mma_b[32][32];
static inline void dmma8x8x4(double& d0, double& d1, double a, double b) {
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    emu_mma_a[w][lane] = a;
    emu_mma_b[w][lane] = b;
    __syncwarp();
    const int row = lane >> 2, col = 2 * (lane & 3);
    for (int k = 0; k < 4; ++k) {
        const double av = emu_mma_a[w][row * 4 + k];
        d0 = std::fma(av, emu_mma_b[w][col * 4 + k], d0);
        d1 = std::fma(av, emu_mma_b[w][(col + 1) * 4 + k], d1);
    }
    __syncwarp();
}
static inline int cudaGetLastError() { return 0; }
static inline unsigned atomicAdd(unsigned* p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
template <class T>
static inline T __ldcg(const T* p) { return *p; }
#define POMS_HIDDEN
// round-to-nearest arithmetic that the compiler must not contract into an FMA
static inline double __dmul_rn(double a, double b) { volatile double r = a * b; return r; }
static inline double __dadd_rn(double a, double b) { volatile double r = a + b; return r; }
using std::max;
using std::min;

// cp.async (8 bytes per thread): the copy happens at once, commit / wait are no-ops -- a missing wait is
// therefore NOT something this emulation can see
static inline void cp_async8(void* smem, const void* gmem) { memcpy(smem, gmem, 8); }
static inline void cp_async_commit() {}
template <int N>
static inline void cp_async_wait() {}
[[maybe_unused]] static int bad_arg(int idx, const char* what) {
    snprintf(g_err, sizeof(g_err), "bad argument %d: %s", idx, what);
    return -idx;
}

#define CHECK_LAUNCH(where) \
    do {                    \
        g_launches++;       \
    } while (0)
#define POMS_LAUNCH(kernel, grid, stream, arg) emu_launch(grid, [&] { kernel(arg); })

// launch with an explicit block size and argument list (sources rewritten by make_emu_source.py)
#ifndef EMU_SMEM_SHRINK
#define EMU_SMEM_SHRINK 0     // self-test of the poisoning: -DEMU_SMEM_SHRINK=64 must make a kernel that uses
#endif                        // all of its dynamic shared memory fail under the address sanitizer
#define EMU_LAUNCH_EX(kernel, grid, block, smem, stream, ...) \
    emu_launch(dim3(grid), [&] { kernel(__VA_ARGS__); }, block, (size_t)(smem) - ((smem) ? EMU_SMEM_SHRINK : 0))

constexpr int EMU_BLOCK = 256;
template <class F>
static void emu_launch(dim3 grid, F body, int nthreads = EMU_BLOCK, size_t dyn_smem = 0) {
    const int EMU_BLOCK = nthreads;               // (shadows the default: the code below is unchanged)
#if defined(__SANITIZE_ADDRESS__)
    // dynamic shared memory: exactly the bytes of this launch are addressable
    const size_t live = (dyn_smem + 7) & ~(size_t)7;
    __asan_unpoison_memory_region(emu_dyn_smem, sizeof(emu_dyn_smem));
    if (live < sizeof(emu_dyn_smem)) __asan_poison_memory_region(emu_dyn_smem + live, sizeof(emu_dyn_smem) - live);
#else
    (void)dyn_smem;
#endif
    blockDim = dim3(nthreads);
    gridDim = grid;
    // a warp whose threads all leave the kernel early must not block the others: the block barrier
    // is only used by kernels that keep every thread (checked by the kernels' own structure)
    for (int w = 0; w < (nthreads + 31) / 32; ++w)
        pthread_barrier_init(&emu_warp_bar[w], nullptr, std::min(32, nthreads - 32 * w));
    pthread_barrier_init(&emu_bar, nullptr, EMU_BLOCK);
    std::vector<std::thread> th;
    th.reserve(EMU_BLOCK);
    for (int t = 0; t < EMU_BLOCK; ++t)
        th.emplace_back([&, t] {
            threadIdx = {(unsigned)t, 0u, 0u};
            for (unsigned z = 0; z < grid.z; ++z)
                for (unsigned y = 0; y < grid.y; ++y)
                    for (unsigned x = 0; x < grid.x; ++x) {
                        blockIdx = {x, y, z};
                        body();
                        pthread_barrier_wait(&emu_bar);   // the next block reuses the "shared" statics
                    }
        });
    for (auto& x : th) x.join();
    pthread_barrier_destroy(&emu_bar);
    for (int w = 0; w < (nthreads + 31) / 32; ++w) pthread_barrier_destroy(&emu_warp_bar[w]);
}
