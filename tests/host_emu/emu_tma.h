// emu_tma.h -- TEST INFRASTRUCTURE ONLY (see cuda_emu.h): host stand-ins for the pieces of the TMA /
// mbarrier pipeline that the 3-D mat-vec kernels reach through inline PTX or the driver API.
//   * CUtensorMap: the real opaque 128-byte struct of <cuda.h>, filled by emu_encode_tiled (handed out by
//     a stub of cudaGetDriverEntryPoint) with base pointer, extents, byte strides and box.
//   * tma_load_3d: copies the box at once (zero fill outside the tensor, like the hardware), then
//     completes its bytes on the mbarrier.
//   * mbarrier: {pending arrivals, outstanding bytes, phase} in the 8-byte shared word, every operation
//     under one mutex (which also gives ThreadSanitizer the acquire / release edges of the real thing);
//     a phase completes when both counts reach zero; mbar_wait spins with sched_yield.
#pragma once
#include <cuda.h>
#include <sched.h>

#include <mutex>

struct EmuTmap {
    const double* base;
    uint64_t dims[3];      // innermost first (n3, n2, n1)
    uint64_t strides[2];   // bytes: row, plane
    uint32_t box[3];
};
static_assert(sizeof(EmuTmap) <= sizeof(CUtensorMap), "descriptor does not fit");

static CUresult emu_encode_tiled(CUtensorMap* out, CUtensorMapDataType, cuuint32_t rank, void* base,
                                 const cuuint64_t* dims, const cuuint64_t* strides, const cuuint32_t* box,
                                 const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                 CUtensorMapFloatOOBfill) {
    if (rank != 2 && rank != 3) return CUDA_ERROR_INVALID_VALUE;
    // the constraints the driver enforces and the kernels rely on
    if (((uintptr_t)base & 15) || (strides[0] & 15) || (rank == 3 && (strides[1] & 15)) || box[0] > 256 || box[1] > 256)
        return CUDA_ERROR_INVALID_VALUE;
    EmuTmap t;
    t.base = (const double*)base;
    for (int i = 0; i < 3; ++i) {
        t.dims[i] = i < (int)rank ? dims[i] : 1;
        t.box[i] = i < (int)rank ? box[i] : 1;
    }
    t.strides[0] = strides[0];
    t.strides[1] = rank == 3 ? strides[1] : 0;
    memset(out, 0, sizeof(*out));
    memcpy(out, &t, sizeof(t));
    return CUDA_SUCCESS;
}
enum cudaDriverEntryPointQueryResult { cudaDriverEntryPointSuccess = 0 };
constexpr int cudaEnableDefault = 0;
static inline int cudaGetDriverEntryPoint(const char*, void** p, int, cudaDriverEntryPointQueryResult* q) {
    *p = (void*)&emu_encode_tiled;
    *q = cudaDriverEntryPointSuccess;
    return cudaSuccess;
}
[[maybe_unused]] static int fail_cuda(cudaError_t e, const char* where) {
    snprintf(g_err, sizeof(g_err), "%s: error %d", where, e);
    return e;
}

struct EmuBar {
    int32_t tx;
    int16_t pending;
    int8_t count;
    int8_t phase;
};
static_assert(sizeof(EmuBar) == 8, "mbarrier word");
static std::mutex emu_bar_mu;
static inline void emu_bar_check(EmuBar* b) {
    if (b->pending == 0 && b->tx == 0) {
        b->phase ^= 1;
        b->pending = b->count;
    }
}
static inline void mbar_init(uint64_t* bar, unsigned count) {
    std::lock_guard<std::mutex> lk(emu_bar_mu);
    EmuBar* b = (EmuBar*)bar;
    b->tx = 0;
    b->pending = (int16_t)count;
    b->count = (int8_t)count;
    b->phase = 0;
}
static inline void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
    std::lock_guard<std::mutex> lk(emu_bar_mu);
    EmuBar* b = (EmuBar*)bar;
    b->tx += (int32_t)bytes;
    b->pending--;
    emu_bar_check(b);
}
static inline void mbar_arrive(uint64_t* bar) {
    std::lock_guard<std::mutex> lk(emu_bar_mu);
    EmuBar* b = (EmuBar*)bar;
    b->pending--;
    emu_bar_check(b);
}
static inline void mbar_wait(uint64_t* bar, unsigned parity) {
    for (;;) {
        {
            std::lock_guard<std::mutex> lk(emu_bar_mu);
            if (((EmuBar*)bar)->phase != (int8_t)parity) return;
        }
        sched_yield();
    }
}
static long long emu_tma_loads;     // bulk tensor copies issued (the harness reports it: TMA path taken?)
static inline void tma_load_3d(void* smem_dst, const CUtensorMap* tmap, int c0, int c1, int c2, uint64_t* bar) {
    __atomic_fetch_add(&emu_tma_loads, 1, __ATOMIC_RELAXED);
    EmuTmap t;
    memcpy(&t, tmap, sizeof(t));
    double* dst = (double*)smem_dst;
    if ((c0 & 1) || ((uintptr_t)dst & 127)) {       // measured constraints
        fprintf(stderr, "emu TMA: misaligned start coordinate %d or destination\n", c0);
        abort();
    }
    for (uint32_t r = 0; r < t.box[1]; ++r)
        for (uint32_t c = 0; c < t.box[0]; ++c) {
            const int64_t g3 = (int64_t)c0 + c, g2 = (int64_t)c1 + r, g1 = c2;
            double v = 0.0;
            if (g3 >= 0 && g3 < (int64_t)t.dims[0] && g2 >= 0 && g2 < (int64_t)t.dims[1] && g1 >= 0 &&
                g1 < (int64_t)t.dims[2])
                v = t.base[g1 * (int64_t)(t.strides[1] / 8) + g2 * (int64_t)(t.strides[0] / 8) + g3];
            dst[(size_t)r * t.box[0] + c] = v;
        }
    std::lock_guard<std::mutex> lk(emu_bar_mu);
    EmuBar* b = (EmuBar*)bar;
    b->tx -= (int32_t)(t.box[0] * t.box[1] * 8);
    emu_bar_check(b);
}

static inline void tma_load_2d(void* smem_dst, const CUtensorMap* tmap, int c0, int c1, uint64_t* bar) {
    tma_load_3d(smem_dst, tmap, c0, c1, 0, bar);       // rank-2 maps are stored with a unit third extent
}

// warp vote through a per-warp exchange buffer
static unsigned emu_vote_buf[32][32];
static inline unsigned __ballot_sync(unsigned, bool pred) {
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    emu_vote_buf[w][lane] = pred ? 1u : 0u;
    __syncwarp();
    unsigned r = 0;
    for (int l = 0; l < 32; ++l) r |= emu_vote_buf[w][l] << l;
    __syncwarp();
    return r;
}
#define __grid_constant__
