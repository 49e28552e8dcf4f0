// emu_halo.cpp -- TEST INFRASTRUCTURE ONLY (see cuda_emu.h): the peer-store halo exchange of poms_extra.cu
// (halo_push_kernel behind poms_halo_exchange_p2p, rewritten for g++ by make_emu_source.py into halo_emu.cuh).
// One PROCESS per rank, like the product (one process per GPU); the "IPC-mapped" memory of all ranks is one file
// mapped MAP_SHARED by every process, so the flag protocol (ENTER / DATA sequence numbers, block ticket) runs
// between real concurrent address spaces.  st.release.sys / ld.acquire.sys are sequentially consistent atomics.
//   emu_halo <shared file> <rank> <size> <n_doubles> <exchanges> <seed>
// shared file layout, per rank r (stride = 64 + 8 * (2 n + 8) bytes; a canary = 2 doubles, keeps 16-byte alignment):
//   flags (8 x uint64) | canary | ghost_lo (n doubles) | canary | canary | ghost_hi (n doubles) | canary
// Exchange s: every rank fills its lowest / highest owned planes with f(rank, s, side, i), sleeps a random time,
// exchanges, sleeps again (a slow consumer of the ghost planes), and checks that its ghost planes hold the neighbours' planes OF EXCHANGE s (a neighbour that ran
// ahead and overwrote them, or data that had not landed, shows as a mismatch).  Exit code 0 = all exchanges exact.
#define POMS_HOST_EMU 1
#include "cuda_emu.h"
#include "poms_b200.h"
#include <fcntl.h>
#include <sched.h>
#include <sys/mman.h>
#include <unistd.h>
#include <cstdlib>
static inline void st_release_sys(uint64_t* p, uint64_t v) { __atomic_store_n(p, v, __ATOMIC_SEQ_CST); }
static inline uint64_t ld_acquire_sys(const uint64_t* p) { return __atomic_load_n(p, __ATOMIC_SEQ_CST); }
static inline void __nanosleep(unsigned) { sched_yield(); }
static inline void __threadfence_system() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
static int x_bad_arg(int idx, const char* what) {
    snprintf(g_err, sizeof(g_err), "bad argument %d: %s", idx, what);
    return -idx;
}
static int x_fail_cuda(cudaError_t e, const char* where) {
    snprintf(g_err, sizeof(g_err), "%s: error %d", where, e);
    return e;
}
#include "halo_emu.cuh"

static const double CANARY = -7.25e300;
static double f(int rank, int s, int side, int64_t i) { return rank * 1e6 + s * 1e3 + side * 0.5 + 1e-3 * (double)(i % 977); }

int main(int argc, char** argv) {
    if (argc != 7) return 2;
    const int rank = atoi(argv[2]), size = atoi(argv[3]), S = atoi(argv[5]);
    const int64_t n = atoll(argv[4]);
    srand(atoi(argv[6]) * 31 + rank);
    const size_t stride = 64 + 8 * (size_t)(2 * n + 8);
    const int fd = open(argv[1], O_RDWR);
    if (fd < 0) return 3;
    unsigned char* base = (unsigned char*)mmap(nullptr, stride * size, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    if (base == MAP_FAILED) return 3;
    auto flags = [&](int r) { return (uint64_t*)(base + stride * r); };
    auto ghost_lo = [&](int r) { return (double*)(base + stride * r + 64) + 2; };
    auto ghost_hi = [&](int r) { return (double*)(base + stride * r + 64) + n + 6; };
    double *src_lo, *src_hi;                                   // owned planes: private memory, exactly sized
    if (posix_memalign((void**)&src_lo, 16, 8 * (size_t)(n ? n : 2)) || posix_memalign((void**)&src_hi, 16, 8 * (size_t)(n ? n : 2))) return 3;
    const bool lo = rank > 0, hi = rank + 1 < size;
    int bad = 0;
    for (int s = 1; s <= S; ++s) {
        for (int64_t i = 0; i < n; ++i) {
            src_lo[i] = f(rank, s, 0, i);
            src_hi[i] = f(rank, s, 1, i);
        }
        usleep(rand() % 3 == 0 ? 60000 : rand() % 3000);       // ranks drift apart between exchanges (now and then by a lot:
                                                               // a neighbour's data arrive long after this rank entered)
        const int rc = poms_halo_exchange_p2p(src_lo, lo ? ghost_hi(rank - 1) : nullptr, src_hi, hi ? ghost_lo(rank + 1) : nullptr,
                                              n, flags(rank), lo ? flags(rank - 1) : nullptr, hi ? flags(rank + 1) : nullptr, nullptr);
        if (rc != 0) {
            fprintf(stderr, "status %d: %s\n", rc, g_err);
            return 4;
        }
        if (rand() % 2) usleep(rand() % 3 == 0 ? 60000 : rand() % 3000);   // now and then a slow consumer: the ghost planes are read late
                                                               // (longer than an emulated kernel start, so a neighbour could run ahead)
        for (int64_t i = 0; i < n; ++i) {
            if (lo && ghost_lo(rank)[i] != f(rank - 1, s, 1, i)) ++bad;
            if (hi && ghost_hi(rank)[i] != f(rank + 1, s, 0, i)) ++bad;
        }
        if (flags(rank)[0] != (uint64_t)s || ((unsigned*)(flags(rank) + 5))[0] != 0u) ++bad;   // sequence advanced, ticket reset
    }
    const double* c = (const double*)(base + stride * rank + 64);
    bool ok = true;
    for (int k = 0; k < 2; ++k) ok = ok && c[k] == CANARY && c[n + 2 + k] == CANARY && c[n + 4 + k] == CANARY && c[2 * n + 6 + k] == CANARY;
    if (!ok) {
        fprintf(stderr, "rank %d: canary overwritten\n", rank);
        return 5;
    }
    if (bad) fprintf(stderr, "rank %d: %d mismatches\n", rank, bad);
    free(src_lo);
    free(src_hi);
    return bad ? 1 : 0;
}
