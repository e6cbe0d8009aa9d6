// common.cuh — shared host/device helpers for libdiskrag_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <atomic>
#include <mutex>
#include <string>

#include "../../include/diskrag_b200.h"

typedef unsigned long long u64;

// ---------------------------------------------------------------------------------------------
// host side: errors, launch accounting, the index handle
// ---------------------------------------------------------------------------------------------
void dr_set_error(const char *fmt, ...);
extern std::atomic<long long> g_launches;

#define DR_CUDA(call)                                                                              \
    do {                                                                                           \
        cudaError_t e__ = (call);                                                                  \
        if (e__ != cudaSuccess) {                                                                  \
            dr_set_error("%s failed at %s:%d: %s", #call, __FILE__, __LINE__, cudaGetErrorString(e__)); \
            return 1;                                                                              \
        }                                                                                          \
    } while (0)

#define DR_CHECK(cond, ...)                \
    do {                                   \
        if (!(cond)) {                     \
            dr_set_error(__VA_ARGS__);     \
            return 2;                      \
        }                                  \
    } while (0)

#define DR_LAUNCHED()                                  \
    do {                                               \
        g_launches.fetch_add(1);                       \
        DR_CUDA(cudaGetLastError());                   \
    } while (0)

#define DR_PIPE_EVENTS 16

struct dr_index {
    int device = 0;
    int64_t N = 0;
    int64_t cap = 0;              // rows allocated (>= N) when the arrays are owned: dr_index_append grows them
    int D = 0, R = 0, M = 0;
    int64_t medoid = 0;
    bool owns = true;
    float *d_vec = nullptr;       // [N, D]
    uint32_t *d_adj = nullptr;    // [N, R] 0-padded, stored order
    int32_t *d_deg = nullptr;     // optional true degrees (only while a graph is being built)
    uint8_t *d_deleted = nullptr; // optional lazy-delete mask (vamana_graph.py:116-125): such nodes are never visited
    uint8_t *d_codes = nullptr;   // [N, M]
    float *d_codebook = nullptr;  // [M, 256, ds]
    // scratch (grown on demand, reused across calls)
    float *d_lut = nullptr; size_t lut_bytes = 0;
    u64 *d_counter = nullptr;
    uint32_t *d_ovf = nullptr; size_t ovf_bytes = 0;
    void *d_io = nullptr; size_t io_bytes = 0;  // staging for the host-pointer API
    int sms = 0, smem_optin = 0;
    // search-kernel timing (bench roofline)
    bool timing = false; double timed_ms = 0.0; long long timed_launches = 0;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    // last use of the shared scratch: a search enqueued on a different stream waits for it on the device
    cudaEvent_t ev_scratch = nullptr; cudaStream_t scratch_stream = nullptr; bool scratch_used = false;
    // index-sharded exchange fused into the search kernel's epilogue (dr_index_set_peer_route): every query's top-k goes, as packed
    // keys, straight into the receive buffer of the rank that reduces it (peer memory over NVLink), in addition to the local output
    u64 *const *d_peer_recv = nullptr;  // device array of G pointers: rank g's receive buffer [G][Bq][k]
    int peer_G = 0, peer_rank = 0; int64_t peer_B = 0, peer_id_offset = 0;
    // host-pointer API pipeline (copy-in / compute / copy-out streams)
    cudaStream_t s_in = nullptr, s_comp = nullptr, s_out = nullptr;
    cudaEvent_t ev_in[DR_PIPE_EVENTS] = {}, ev_done[DR_PIPE_EVENTS] = {};
    // The scratch buffers above belong to the handle, so entry points that touch them hold this lock for the duration of the
    // call: concurrent callers (FastAPI worker threads, ctypes releases the GIL) are serialised per handle instead of racing.
    std::mutex mu;
};
#define DR_LOCK(h) std::lock_guard<std::mutex> dr_lock_guard_((h)->mu)

int dr_scratch(void **ptr, size_t *cur, size_t need);  // grow-only device scratch

// internal launchers (defined in the respective .cu files)
int launch_lut_build(const float *d_codebook, const float *d_Q, int64_t B, int D, int M, float *d_out, cudaStream_t s);
int launch_search(dr_index *h, const float *d_Q, int64_t B, const dr_search_params *p, const float *d_lut,
                  int32_t *ids, float *dist, int32_t *hops, int32_t *visited, int32_t *list_ids, float *list_dist,
                  int32_t *list_len, int32_t *trace, int32_t trace_cap, int32_t *status, cudaStream_t s,
                  const int32_t *qmap = nullptr);
int launch_search_fast(dr_index *h, const float *d_Q, int64_t B, const dr_search_params *p, int32_t *ids, float *dist,
                       int32_t *hops, int32_t *visited, int32_t *list_ids, float *list_dist, int32_t *list_len,
                       int32_t *status, cudaStream_t s);   // search_fast.cu: u8-table throughput kernel
void beam_c_plan(const dr_index *h, int64_t B, long long *grid_out, size_t *bitmap_bytes_out);   // beam_c.cu
int launch_beam_c(dr_index *h, const float *d_Q, int64_t B, int k, int bw, int dist, int sqrt_out, const float *d_lut,
                  uint32_t *d_bitmaps, int32_t *ids, float *dists, int32_t *hops, int32_t *visited, cudaStream_t s,
                  int64_t start = -1);
int launch_lut_build_u8(const float *d_codebook, const float *d_Q, int64_t B, int D, int M, uint8_t *d_out8, float *d_scale,
                        float *d_offset, float *d_mn, unsigned *d_range, int word_layout, cudaStream_t s);  // pq.cu
int launch_lut_build_u8_tc(const float *d_codebook, const float *d_Q, int64_t B, int D, int M, uint8_t *d_out8, float *d_scale,
                           float *d_offset, float *d_mn, unsigned *d_range, int sms, cudaStream_t s);   // lut_tc.cu (tcgen05)
void pq_train_set_kmeanspp(int enable);        // pq.cu: k-means++ seeding (default) or evenly spaced rows
void pq_train_set_tensor_cores(int enable);   // pq.cu: k-means assignment on tcgen05 (default) or exact fp32
int launch_lut_u8_unpermute(const uint8_t *d_words, int64_t B, int M, uint8_t *d_plain, cudaStream_t s);  // pq.cu

// ---------------------------------------------------------------------------------------------
// device side
// ---------------------------------------------------------------------------------------------
#ifdef __CUDACC__

#define DR_FULL 0xFFFFFFFFu
#define DR_EMPTY 0xFFFFFFFFu
#define DR_KEY_MAX 0xFFFFFFFFFFFFFFFFull

// order-preserving float -> uint32 (total order equals the float order; -0.0 is folded into +0.0 by the caller)
__device__ __forceinline__ uint32_t f2ord(float f) {
    uint32_t b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float ord2f(uint32_t o) {
    uint32_t b = (o & 0x80000000u) ? (o & 0x7FFFFFFFu) : ~o;
    return __uint_as_float(b);
}
// search-list key: (distance, id) ascending == the reference's tuple order; bit 0 = expanded flag
__device__ __forceinline__ u64 make_key(float d, uint32_t id) {
    return ((u64)f2ord(d + 0.0f) << 32) | ((u64)id << 1);
}
__device__ __forceinline__ uint32_t key_id(u64 k) { return (uint32_t)(k & 0xFFFFFFFFull) >> 1; }
__device__ __forceinline__ float key_dist(u64 k) { return ord2f((uint32_t)(k >> 32)); }
__device__ __forceinline__ uint32_t key_dbits(u64 k) { return (uint32_t)(k >> 32); }

__device__ __forceinline__ float warp_sum_butterfly(float v) {
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) v = __fadd_rn(v, __shfl_xor_sync(DR_FULL, v, off));
    return v;
}

__device__ __forceinline__ float4 ldg_f4(const float *p) { return __ldg(reinterpret_cast<const float4 *>(p)); }

// a - b per component, IEEE round-to-nearest like four __fsub_rn; with DR_F32X2 the four subtractions are two packed
// sub.rn.f32x2 (FADD2 on sm_100a: one issue slot for two results; the vector loads already deliver aligned register pairs)
#ifndef DR_F32X2
#define DR_F32X2 1
#endif
__device__ __forceinline__ void f4sub(const float4 &a, const float4 &b, float &d0, float &d1, float &d2, float &d3) {
#if DR_F32X2 && defined(__CUDA_ARCH__)   // (the host-side warp emulator of tests/ compiles this text with g++: scalar path)
    unsigned long long alo, ahi, blo, bhi, dlo, dhi;
    asm("mov.b64 %0, {%1, %2};" : "=l"(alo) : "f"(a.x), "f"(a.y));
    asm("mov.b64 %0, {%1, %2};" : "=l"(ahi) : "f"(a.z), "f"(a.w));
    asm("mov.b64 %0, {%1, %2};" : "=l"(blo) : "f"(b.x), "f"(b.y));
    asm("mov.b64 %0, {%1, %2};" : "=l"(bhi) : "f"(b.z), "f"(b.w));
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(dlo) : "l"(alo), "l"(blo));
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(dhi) : "l"(ahi), "l"(bhi));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(d0), "=f"(d1) : "l"(dlo));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(d2), "=f"(d3) : "l"(dhi));
#else
    d0 = __fsub_rn(a.x, b.x); d1 = __fsub_rn(a.y, b.y); d2 = __fsub_rn(a.z, b.z); d3 = __fsub_rn(a.w, b.w);
#endif
}

// Canonical exact fp32 squared L2 of one row against a query held in shared memory, one warp per row.
// Order (restated by oracle.c:orc_l2sq_warp): lane l owns elements (j*32+l)*VW+c, VW = 4 if D%4==0
// else 1; per lane fmaf accumulation in increasing index; xor-butterfly 16,8,4,2,1.
__device__ __forceinline__ float warp_l2sq(const float *__restrict__ row, const float *__restrict__ q, int D, int lane) {
    float acc = 0.0f;
    if ((D & 3) == 0) {
        int base = lane * 4;
#pragma unroll 4
        for (; base < D; base += 128) {
            float4 a = ldg_f4(row + base);
            float4 b = *reinterpret_cast<const float4 *>(q + base);
            float d0, d1, d2, d3;
            f4sub(a, b, d0, d1, d2, d3);
            acc = __fmaf_rn(d0, d0, acc);
            acc = __fmaf_rn(d1, d1, acc);
            acc = __fmaf_rn(d2, d2, acc);
            acc = __fmaf_rn(d3, d3, acc);
        }
    } else {
        for (int i = lane; i < D; i += 32) {
            float d = __fsub_rn(__ldg(row + i), q[i]);
            acc = __fmaf_rn(d, d, acc);
        }
    }
    return warp_sum_butterfly(acc);
}

// Canonical cosine DISTANCE 1 - cos (cosine_similarity_cython, cython_utils.pyx:53-70; 0.0 when a norm is 0) of one row
// against the query in shared memory, one warp per row: the three sums (x.q, x.x, q.q) use the lane / fmaf / butterfly
// order of distance.cu:rowdist_kernel (op 2), the final division is done in double like the reference (np.sqrt of a
// Python float).  Restated by oracle.c:orc_cosine_warp.
__device__ __forceinline__ float warp_cosdist(const float *__restrict__ row, const float *__restrict__ q, int D, int lane) {
    float s0 = 0.0f, s1 = 0.0f, s2 = 0.0f;
    if ((D & 3) == 0) {
#pragma unroll 4
        for (int base = lane * 4; base < D; base += 128) {
            const float4 x = ldg_f4(row + base);
            const float4 y = *reinterpret_cast<const float4 *>(q + base);
            const float xs[4] = {x.x, x.y, x.z, x.w}, ys[4] = {y.x, y.y, y.z, y.w};
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                s0 = __fmaf_rn(xs[c], ys[c], s0); s1 = __fmaf_rn(xs[c], xs[c], s1); s2 = __fmaf_rn(ys[c], ys[c], s2);
            }
        }
    } else {
        for (int i = lane; i < D; i += 32) {
            const float x = __ldg(row + i), y = q[i];
            s0 = __fmaf_rn(x, y, s0); s1 = __fmaf_rn(x, x, s1); s2 = __fmaf_rn(y, y, s2);
        }
    }
    s0 = warp_sum_butterfly(s0); s1 = warp_sum_butterfly(s1); s2 = warp_sum_butterfly(s2);
    if (s1 == 0.0f || s2 == 0.0f) return 0.0f;
    return (float)(1.0 - ((double)s0 / (sqrt((double)s1) * sqrt((double)s2))));
}

// two rows at once (twice the loads in flight per lane); each row keeps the canonical order of warp_l2sq
__device__ __forceinline__ void warp_l2sq_x2(const float *__restrict__ rowA, const float *__restrict__ rowB,
                                             const float *__restrict__ q, int D, int lane, float &dA, float &dB) {
    float accA = 0.0f, accB = 0.0f;
    if ((D & 3) == 0) {
        int base = lane * 4;
#pragma unroll 4
        for (; base < D; base += 128) {
            float4 a = ldg_f4(rowA + base);
            float4 c = ldg_f4(rowB + base);
            float4 b = *reinterpret_cast<const float4 *>(q + base);
            float d0, d1, d2, d3;
            f4sub(a, b, d0, d1, d2, d3);
            accA = __fmaf_rn(d0, d0, accA); accA = __fmaf_rn(d1, d1, accA); accA = __fmaf_rn(d2, d2, accA); accA = __fmaf_rn(d3, d3, accA);
            float e0, e1, e2, e3;
            f4sub(c, b, e0, e1, e2, e3);
            accB = __fmaf_rn(e0, e0, accB); accB = __fmaf_rn(e1, e1, accB); accB = __fmaf_rn(e2, e2, accB); accB = __fmaf_rn(e3, e3, accB);
        }
    } else {
        for (int i = lane; i < D; i += 32) {
            float qq = q[i];
            float d = __fsub_rn(__ldg(rowA + i), qq), e = __fsub_rn(__ldg(rowB + i), qq);
            accA = __fmaf_rn(d, d, accA);
            accB = __fmaf_rn(e, e, accB);
        }
    }
    dA = warp_sum_butterfly(accA);
    dB = warp_sum_butterfly(accB);
}

// ---- mbarrier + bulk async copy (TMA 1-D) helpers ---------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {}
}
// global -> shared bulk copy, completion signalled on an mbarrier (bytes % 16 == 0, both sides 16 B aligned)
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// the same with an L2 eviction-priority hint (createpolicy): streaming rows that are read once should not displace
// the graph / code rows other CTAs are about to gather
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void bulk_g2s_hint(void *dst, const void *src, uint32_t bytes, uint64_t *bar, uint64_t pol) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(pol) : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

__device__ __forceinline__ uint32_t hash_u32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}

// visited set: open-addressing table in shared memory (atomicCAS claims a slot); once it is 3/4 full the shared table
// is frozen (lookups only) and new ids go to the CTA's global overflow table.  Returns true when nb is first seen.
__device__ __forceinline__ bool visited_insert(uint32_t nb, uint32_t *hash, uint32_t mask, bool use_ovf,
                                               uint32_t *ovf, uint32_t ovf_mask) {
    uint32_t h = hash_u32(nb) & mask;
    if (!use_ovf) {
        for (;;) {
            uint32_t old = atomicCAS(&hash[h], DR_EMPTY, nb);
            if (old == DR_EMPTY) return true;
            if (old == nb) return false;
            h = (h + 1) & mask;
        }
    }
    for (;;) {  // shared table is frozen: look up only, then claim in the global overflow table
        uint32_t cur = hash[h];
        if (cur == nb) return false;
        if (cur == DR_EMPTY) break;
        h = (h + 1) & mask;
    }
    h = hash_u32(nb ^ 0x9e3779b9u) & ovf_mask;
    for (;;) {
        uint32_t old = atomicCAS(&ovf[h], DR_EMPTY, nb);
        if (old == DR_EMPTY) return true;
        if (old == nb) return false;
        h = (h + 1) & ovf_mask;
    }
}

#endif  // __CUDACC__
