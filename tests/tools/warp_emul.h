// warp_emul.h — TEST INFRASTRUCTURE: runs a one-warp CUDA kernel on the CPU, one OS thread per lane, in lockstep.
// Every warp collective the kernel uses (__syncwarp, __shfl_sync, __shfl_xor_sync, __ballot_sync, __match_any_sync) is an exchange
// through a 32-slot board between two barriers, which is exact as long as the kernel calls them in warp-uniform control flow (it has to,
// on the GPU too: they are all issued with the full mask).  Atomics map to the compiler's; the rounding-mode intrinsics to plain
// IEEE operations (compile with -ffp-contract=off).  Used by tests/test_beam_c_emulated.py to execute csrc/beam_c.cu's kernel text —
// extracted from the .cu file at test time, not copied — against the oracle where no GPU exists.
#pragma once
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
using std::max;
using std::min;

typedef unsigned long long u64;
#define __device__
#define __global__
#define __forceinline__ inline
#define __noinline__
#define __restrict__
#define __launch_bounds__(...)
#define __align__(x)
#define __shared__
#define DR_DIST_PQ 0
#define DR_DIST_EXACT 1

struct emul_dim { unsigned x; };
static thread_local emul_dim threadIdx;
static emul_dim blockIdx = {0}, gridDim = {1};
struct float4 { float x, y, z, w; };

static pthread_barrier_t emul_bar;
static volatile uint64_t emul_slot[32];
static inline void emul_sync() { pthread_barrier_wait(&emul_bar); }
template <class T> static inline T emul_shfl(T v, int src) {
    uint64_t raw = 0;
    memcpy(&raw, &v, sizeof(T));
    emul_slot[threadIdx.x] = raw;
    emul_sync();
    raw = emul_slot[src & 31];
    emul_sync();
    T r;
    memcpy(&r, &raw, sizeof(T));
    return r;
}
static inline unsigned emul_ballot(bool p) {
    emul_slot[threadIdx.x] = p ? 1 : 0;
    emul_sync();
    unsigned r = 0;
    for (int i = 0; i < 32; ++i) r |= (unsigned)emul_slot[i] << i;
    emul_sync();
    return r;
}
static inline unsigned emul_match_any(uint32_t v) {
    emul_slot[threadIdx.x] = v;
    emul_sync();
    unsigned r = 0;
    for (int i = 0; i < 32; ++i) r |= (unsigned)(emul_slot[i] == v) << i;
    emul_sync();
    return r;
}
#define __syncwarp() emul_sync()
#define __shfl_sync(m, v, src) emul_shfl(v, src)
static inline int emul_lane() { return (int)threadIdx.x; }
#define __shfl_xor_sync(m_, v_, o_) emul_shfl(v_, emul_lane() ^ (o_))
#define __ballot_sync(m, p) emul_ballot(p)
#define __match_any_sync(m, v) emul_match_any(v)
static inline unsigned atomicOr(uint32_t *p, uint32_t v) { return __atomic_fetch_or(p, v, __ATOMIC_SEQ_CST); }
template <class T> static inline T __ldg(const T *p) { return *p; }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
static inline float __fsqrt_rn(float a) { return sqrtf(a); }
static inline uint32_t __float_as_uint(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float __uint_as_float(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static inline float __int_as_float(int u) { float f; memcpy(&f, &u, 4); return f; }
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline int __ffs(unsigned x) { return __builtin_ffs((int)x); }
