// search_fast.cu — K1f: the throughput form of the batched Vamana beam search (same path as search.cu:
// greedy_search_cython + compute_query_distance, cython_utils.pyx:72-122 / vamana_graph.py:301-329, followed by
// the exact rerank of search_engine.py:374-379), one CTA per query, W expansions per step.
//
// What differs from the reference-order kernel (search.cu):
//   * the ADC table is 8-bit (pq.cu: launch_lut_build_u8): M x 256 BYTES (48 KB at M = 192 instead of 192 KB),
//     so three CTAs share an SM; a distance is an exact integer sum of table bytes, hence independent of the
//     summation order; keys are (sum, id) and unique, so the merge is a pure key-order merge.
//   * shared-memory table layout is centroid-major, one 32-bit word = the entries of 4 consecutive subspaces for
//     one centroid: lane l owns code word l of a row and its four lookups all fall into bank l whatever the code
//     bytes are (no bank conflicts on the first 128 subspaces of a row; the tail words of two rows share a pass).
//   * code words of 4 rows are fetched before any lookup (memory-level parallelism), survivors are compacted,
//     small survivor sets are merged by rank counting, larger ones by per-warp bitonic sorts + binary-search ranks.
// Restated bit-for-bit by oracle.c (dist_mode 4, orc_search_list with strict_ties = 0).
#include "common.cuh"
#include <stdlib.h>

extern __shared__ __align__(16) unsigned char dr_smem[];

struct FastArgs {
    const float *vec; const uint32_t *adj; const uint8_t *codes; const uint8_t *deleted;
    const float *Q; const uint8_t *lut8; const float *lut_scale; const float *lut_offset;
    long long N; int D, R, M;
    long long B; int k, L, W;
    int W2;         // expansions of a step that follows a step without survivors (>= W; == W: off)
    int rerank, sqrt_out, prefetch;
    uint32_t start;
    int32_t *out_ids; float *out_dist; int32_t *out_hops; int32_t *out_visited;
    int32_t *list_ids; float *list_dist; int32_t *list_len; int32_t *status;
    u64 *counter;
    uint32_t *ovf; uint32_t ovf_cap; uint32_t hash_cap;
    int o_q, o_rrk, o_list0, o_list1, o_ur0, o_ur1, o_newk, o_newid, o_sel, o_adjrow, o_hash;
    int rr_slots;   // rerank staging slots available in the table region (the query vector may occupy the last one)
    // index-sharded exchange in the epilogue: rank g = owner of global query qb reduces it; its buffer is [G][Bq][k] packed keys
    u64 *const *peer_recv; int peer_G, peer_rank; long long peer_B, peer_id_offset, peer_q0;
};

// where global query qb of a peer_B-query batch goes: owner rank (contiguous slices, the first peer_B % G ranks hold one more) and
// the row inside the owner's slice
__device__ __forceinline__ void peer_slot(long long qb, long long B, int G, int &g, long long &r) {
    const long long per = B / G, extra = B % G, cut = extra * (per + 1);
    if (qb < cut) { g = (int)(qb / (per + 1)); r = qb - (long long)g * (per + 1); }
    else { g = (int)(extra + (qb - cut) / (per > 0 ? per : 1)); r = qb - cut - (long long)(g - extra) * per; }
}

__device__ __forceinline__ u64 make_ikey(uint32_t sum, uint32_t id) { return ((u64)sum << 32) | ((u64)id << 1); }

// ---- byte path (M % 4 != 0): table [m][c], one row per warp -------------------------------------------
__device__ __forceinline__ uint32_t adc_u8_warp(const uint8_t *__restrict__ code, const uint8_t *__restrict__ lut8, int M, int lane) {
    uint32_t acc = 0u;
    for (int m = lane; m < M; m += 32) acc += lut8[(m << 8) + __ldg(code + m)];
    return __reduce_add_sync(DR_FULL, acc);
}

// ---- word path ---------------------------------------------------------------------------------------
// Shared-memory layout of the table for words = M / 4 code words per row, nfull = words / 32, rem = words % 32:
//   entry (word w, centroid c, byte j) = subspace 4w + j:
//     w <  32 nfull : (w / 32) * 32768 + c * 128     + (w % 32) * 4      + j      (bank = w % 32)
//     w >= 32 nfull : nfull * 32768    + c * 4 * rem + (w - 32 nfull) * 4 + j
// tb points at the lane's word of centroid 0; the four entries selected by the code word's bytes are summed.
// tb = 32-bit shared-memory address of the lane's word of centroid 0.
__device__ __forceinline__ uint32_t lds_u8(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
template <int J>
__device__ __forceinline__ uint32_t lds_u8_off(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1+%2];" : "=r"(v) : "r"(addr), "n"(J));
    return v;
}
// One byte-wise multiply-add forms each lookup address: dp4a(w, stride in byte k, tb) = tb + stride * code byte k (IDP.4A), so a
// lookup is address + LDS.U8 (+ its share of the adds) instead of extract (PRMT) + scale-add (IMAD) + LDS.U8.  stride <= 255.
#ifndef DR_DP4A_ADDR
#define DR_DP4A_ADDR 1
#endif
template <int STRIDE>
__device__ __forceinline__ uint32_t tab_sum4(uint32_t tb, uint32_t w, int stride_rt) {
#if DR_DP4A_ADDR
    if (STRIDE == 128) {
        return lds_u8_off<0>(__dp4a(w, 0x00000080u, tb)) + lds_u8_off<1>(__dp4a(w, 0x00008000u, tb)) +
               lds_u8_off<2>(__dp4a(w, 0x00800000u, tb)) + lds_u8_off<3>(__dp4a(w, 0x80000000u, tb));
    } else {
        const uint32_t st = (uint32_t)stride_rt;     // <= 64: the tail chunk holds at most 16 words per centroid
        return lds_u8_off<0>(__dp4a(w, st, tb)) + lds_u8_off<1>(__dp4a(w, st << 8, tb)) + lds_u8_off<2>(__dp4a(w, st << 16, tb)) +
               lds_u8_off<3>(__dp4a(w, st << 24, tb));
    }
#else
    const uint32_t c0 = __byte_perm(w, 0u, 0x4440), c1 = __byte_perm(w, 0u, 0x4441), c2 = __byte_perm(w, 0u, 0x4442),
                   c3 = __byte_perm(w, 0u, 0x4443);
    if (STRIDE == 128) {
        return lds_u8_off<0>(tb + (c0 << 7)) + lds_u8_off<1>(tb + (c1 << 7)) + lds_u8_off<2>(tb + (c2 << 7)) +
               lds_u8_off<3>(tb + (c3 << 7));
    } else {
        const uint32_t st = (uint32_t)stride_rt;
        return lds_u8_off<0>(tb + c0 * st) + lds_u8_off<1>(tb + c1 * st) + lds_u8_off<2>(tb + c2 * st) + lds_u8_off<3>(tb + c3 * st);
    }
#endif
}

// G code rows per warp (G even): lane l owns word 32 k + l of every full chunk; the tail words (rem <= 16) of TWO
// rows share one pass (lanes 0-15 row 2p, lanes 16-31 row 2p+1).  Loading and summing are separate so that the next
// group's code words are in flight while this group goes through the table.
// Every ids[g] must be a valid row (callers pad a short group by repeating an id).
template <int G>
struct RowWords { uint32_t wf[2][G]; uint32_t wt[G]; };

// base_lane = codes + 4 * lane (word `lane` of row 0), base_tail = codes + 128 * nfull + 4 * (lane & 15): the callers keep both in
// registers (one 32 x 32 + 64-bit multiply-add per row address instead of a 64-bit add chain on a uniform base)
template <int WORDS, int G>
__device__ __forceinline__ void rows_load(const uint8_t *__restrict__ base_lane, const uint8_t *__restrict__ base_tail, int M,
                                          const uint32_t (&ids)[G], int lane, RowWords<G> &w) {
    const int words = WORDS ? WORDS : (M >> 2);
    const uint32_t Mc = WORDS ? (uint32_t)(WORDS * 4) : (uint32_t)M;   // row stride in bytes
    const int nfull = words >> 5, rem = words & 31;
    const bool pair = rem > 0 && rem <= 16;
    const int hl = lane & 15;
#pragma unroll
    for (int g = 0; g < G; ++g) {
        const uint32_t *rp = reinterpret_cast<const uint32_t *>(base_lane + (unsigned long long)ids[g] * Mc);
#pragma unroll
        for (int k = 0; k < 2; ++k)
            if (k < nfull) w.wf[k][g] = __ldg(rp + 32 * k);
    }
    if (pair) {
#pragma unroll
        for (int p = 0; p < G / 2; ++p) {
            const uint32_t id = (lane < 16) ? ids[2 * p] : ids[2 * p + 1];
            w.wt[p] = (hl < rem) ? __ldg(reinterpret_cast<const uint32_t *>(base_tail + (unsigned long long)id * Mc)) : 0u;
        }
    } else if (rem) {
#pragma unroll
        for (int g = 0; g < G; ++g)
            w.wt[g] = (lane < rem) ? __ldg(reinterpret_cast<const uint32_t *>(base_lane + (unsigned long long)ids[g] * Mc) + 32 * nfull) : 0u;
    }
}

// sums of the first 2 * npairs rows (npairs is warp-uniform, 1 <= npairs <= G / 2); the other out[] are 0
template <int WORDS, int G>
__device__ __forceinline__ void rows_sum(int M, uint32_t tab, const RowWords<G> &w, int npairs, int lane,
                                         uint32_t (&out)[G]) {
    const int words = WORDS ? WORDS : (M >> 2);
    const int nfull = words >> 5, rem = words & 31;
    const bool pair = rem > 0 && rem <= 16;
    const int hl = lane & 15;
    const int st = rem * 4;
#pragma unroll
    for (int p = 0; p < G / 2; ++p) {
        out[2 * p] = 0u; out[2 * p + 1] = 0u;
        if (p < npairs) {
            uint32_t a0 = 0u, a1 = 0u;
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                if (k < nfull) {
                    const uint32_t tb = tab + k * 32768 + lane * 4;
                    a0 += tab_sum4<128>(tb, w.wf[k][2 * p], 128);
                    a1 += tab_sum4<128>(tb, w.wf[k][2 * p + 1], 128);
                }
            }
            if (pair) {
                const uint32_t tb = tab + nfull * 32768 + hl * 4;
                const uint32_t u = (hl < rem) ? tab_sum4<0>(tb, w.wt[p], st) : 0u;
                a0 += (lane < 16) ? u : 0u;
                a1 += (lane < 16) ? 0u : u;
            } else if (rem) {
                const uint32_t tb = tab + nfull * 32768 + lane * 4;
                a0 += (lane < rem) ? tab_sum4<0>(tb, w.wt[2 * p], st) : 0u;
                a1 += (lane < rem) ? tab_sum4<0>(tb, w.wt[2 * p + 1], st) : 0u;
            }
            out[2 * p] = __reduce_add_sync(DR_FULL, a0);
            out[2 * p + 1] = __reduce_add_sync(DR_FULL, a1);
        }
    }
}

template <int WORDS, int G>
__device__ __forceinline__ void adc_u8_rows(const uint8_t *__restrict__ codes, int M, const uint8_t *__restrict__ tab,
                                            const uint32_t (&ids)[G], int lane, uint32_t (&out)[G]) {
    RowWords<G> w;
    const int words_ = WORDS ? WORDS : (M >> 2);
    rows_load<WORDS, G>(codes + lane * 4, codes + 128 * (words_ >> 5) + (lane & 15) * 4, M, ids, lane, w);
    rows_sum<WORDS, G>(M, smem_u32(tab), w, G / 2, lane, out);
}

// visited set of this kernel: same open-addressing scheme as common.cuh:visited_insert, multiplicative (Fibonacci)
// hash: the slot is the top log2(cap) bits of id * 2654435761
__device__ __forceinline__ uint32_t fib_slot(uint32_t id, uint32_t shift) { return (id * 2654435761u) >> shift; }
__device__ __forceinline__ bool visited_insert_fast(uint32_t nb, uint32_t *hash, uint32_t mask, uint32_t shift, bool use_ovf,
                                                    uint32_t *ovf, uint32_t ovf_mask, uint32_t ovf_shift, bool no_smem) {
    uint32_t h;
    if (no_smem) goto global_table;   // the whole visited set lives in the CTA's global (L2-resident) table
    h = fib_slot(nb, shift);
    if (!use_ovf) {
        for (;;) {
            const uint32_t old = atomicCAS(&hash[h], DR_EMPTY, nb);
            if (old == DR_EMPTY) return true;
            if (old == nb) return false;
            h = (h + 1) & mask;
        }
    }
    for (;;) {  // shared table is frozen: look up only, then claim in the global overflow table
        const uint32_t cur = hash[h];
        if (cur == nb) return false;
        if (cur == DR_EMPTY) break;
        h = (h + 1) & mask;
    }
global_table:
    h = fib_slot(nb ^ 0x9e3779b9u, ovf_shift);
    for (;;) {
        const uint32_t old = atomicCAS(&ovf[h], DR_EMPTY, nb);
        if (old == DR_EMPTY) return true;
        if (old == nb) return false;
        h = (h + 1) & ovf_mask;
    }
}

#ifndef DR_RR_UNROLL
#define DR_RR_UNROLL 2
#endif
#ifndef DR_RR_FULL
#define DR_RR_FULL 6   // D = 1536 specialisations: each half-row piece is 768 elements = 6 float4 per lane, fully unrolled (0 = loop)
#endif
// Rerank rows beyond the staging slots: 0 = fetched only when a slot frees up (48 KB in flight per CTA); 1 = all of them are
// started on their trip to L2 (cp.async.bulk.prefetch.L2, one instruction per 6 KB row) as soon as the traversal ends, so only
// the first round of staged copies waits for DRAM; N >= 2 = a rolling window of N rounds ahead of the staged copies
#ifndef DR_RR_PF
#define DR_RR_PF 0
#endif
__device__ __forceinline__ void bulk_prefetch_l2(const void *src, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
// One piece [e0, e1) of the canonical warp L2^2 (common.cuh:warp_l2sq: lane l owns elements base = 4 l + 128 j, fmaf in
// increasing j), the row read from shared memory where a bulk copy staged it.  e0 is a multiple of 128.
// FULL > 0: the piece is exactly FULL iterations (compile-time dimension): fully unrolled
template <int FULL>
__device__ __forceinline__ float l2sq_piece_smem(const float *__restrict__ row, const float *__restrict__ q, int e0, int e1,
                                                 int lane, float acc) {
    if (FULL > 0) {
        const float *r0 = row + e0 + lane * 4, *q0 = q + e0 + lane * 4;
#pragma unroll
        for (int it = 0; it < FULL; ++it) {
            const float4 x = *reinterpret_cast<const float4 *>(r0 + it * 128);
            const float4 y = *reinterpret_cast<const float4 *>(q0 + it * 128);
            float d0, d1, d2, d3;
            f4sub(x, y, d0, d1, d2, d3);
            acc = __fmaf_rn(d0, d0, acc);
            acc = __fmaf_rn(d1, d1, acc);
            acc = __fmaf_rn(d2, d2, acc);
            acc = __fmaf_rn(d3, d3, acc);
        }
        return acc;
    }
    constexpr int kUnroll = DR_RR_UNROLL;
#pragma unroll kUnroll
    for (int base = e0 + lane * 4; base < e1; base += 128) {
        const float4 x = *reinterpret_cast<const float4 *>(row + base);
        const float4 y = *reinterpret_cast<const float4 *>(q + base);
        float d0, d1, d2, d3;
        f4sub(x, y, d0, d1, d2, d3);
        acc = __fmaf_rn(d0, d0, acc);
        acc = __fmaf_rn(d1, d1, acc);
        acc = __fmaf_rn(d2, d2, acc);
        acc = __fmaf_rn(d3, d3, acc);
    }
    return acc;
}

// number of keys of the sorted sequence a[0..n) that are smaller than key (n < 2 * TOP)
#ifndef DR_BSEARCH_BF
#define DR_BSEARCH_BF 1   // 1: fixed-trip power-of-two descent (no data-dependent loop, no divergence between lanes): +1.5 % QPS; 0: classic loop
#endif
template <int TOP>
__device__ __forceinline__ int lower_bound_u64(const u64 *a, int n, u64 key) {
#if DR_BSEARCH_BF
    int pos = 0;
#pragma unroll
    for (int step = TOP; step >= 1; step >>= 1) {
        const int p = pos + step;
        if (p <= n && a[p - 1] < key) pos = p;
    }
    return pos;
#else
    int lo = 0, hi = n;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (a[mid] < key) lo = mid + 1; else hi = mid;
    }
    return lo;
#endif
}

// optional per-phase cycle accounting (build with -DDR_PHASE_TIMING; thread 0 of every CTA, summed into counter[1..8])
#ifdef DR_PHASE_TIMING
#define DR_PT(i) do { if (tid == 0) { const long long c__ = clock64(); pt_acc[i] += c__ - pt_last; pt_last = c__; } } while (0)
#else
#define DR_PT(i) do { } while (0)
#endif

#ifndef DR_FAST_NT
#define DR_FAST_NT 256   // threads per CTA of the throughput kernel (three CTAs per SM)
#endif
#ifndef DR_MERGE_LINEAR
#define DR_MERGE_LINEAR 64   // 32 -> 64: +0.8 % (the one-warp chunk sort of 33..64 survivors kept seven warps at a barrier)
#endif  // up to this many survivors: rank by counting, no sort
// prefetch mask the serving-shape specialisation (RW8 = 4) is compiled for (dr_search_params.prefetch must equal it)
#ifndef DR_PF_SPEC
#define DR_PF_SPEC 5
#endif
// "empty-step doubling" of the serving-shape specialisations (RW8 >= 1 fixes W = 8): dr_search_params.w_after_empty they are compiled for
#ifndef DR_W2_SPEC
#define DR_W2_SPEC 20
#endif
#ifndef DR_MERGE_PAIRS
#define DR_MERGE_PAIRS 0   // experiment: linear merge with two lanes per item (all eight warps busy, half the scan per item): -0.3 % (more
                           // instructions issued for a shorter critical path: the kernel is issue-bound, not latency-bound, there)
#endif
#ifndef DR_SELCAP
#define DR_SELCAP 64    // ranks recorded at merge time (>= W + 3 * W2 covers four steps in a row without a survivor)
#endif
#ifndef DR_L2V
#define DR_L2V 0   // experiment (scripts/build_variants.py): 1 = the serving-shape specialisations keep the visited set in the
#endif             // CTA's L2-resident table (no shared-memory hash) so that four CTAs fit on an SM; pair with DR_FAST_NT=192

// WORDS > 0: compile-time M / 4;  WORDS == 0: runtime M (M % 4 == 0, M <= 256);  WORDS < 0: byte path (any M)
// MINB = CTAs per SM the register budget is cut for (3: 80 registers; 4: 64 registers, used when the visited set moves out
// of shared memory so that a fourth CTA fits)
// RW8 = 1: R == 32 and W == 8 are compile-time; RW8 = 2: also D == 1536; RW8 = 3: also L == 100 and a 4096-slot visited table
// (the bench / default serving shape: text-embedding-3-small vectors, beam 100); RW8 = 4: also prefetch mask 5, no
// lazy-delete mask, rerank on
template <int WORDS, int MINB, int RW8>
__global__ void __launch_bounds__(DR_FAST_NT, MINB) search_fast_kernel(const FastArgs a) {
    constexpr bool WP = WORDS >= 0;
    uint8_t *s_lut = dr_smem;
    float *s_q = reinterpret_cast<float *>(dr_smem + a.o_q);
    u64 *s_list0 = reinterpret_cast<u64 *>(dr_smem + a.o_list0);
    u64 *s_list1 = reinterpret_cast<u64 *>(dr_smem + a.o_list1);
    uint16_t *s_ur0 = reinterpret_cast<uint16_t *>(dr_smem + a.o_ur0);   // per list entry: unexpanded entries before it
    uint16_t *s_ur1 = reinterpret_cast<uint16_t *>(dr_smem + a.o_ur1);
    u64 *s_newk = reinterpret_cast<u64 *>(dr_smem + a.o_newk);
    uint32_t *s_newid = reinterpret_cast<uint32_t *>(dr_smem + a.o_newid);
    uint32_t *s_sel = reinterpret_cast<uint32_t *>(dr_smem + a.o_sel);      // [0, W): next-in-line positions; [W, 2W): the selection
    uint32_t *s_adjrow = reinterpret_cast<uint32_t *>(dr_smem + a.o_adjrow);   // prefetch & 16: the W selected nodes' adjacency rows
    uint32_t *s_hash = reinterpret_cast<uint32_t *>(dr_smem + a.o_hash);
    u64 *s_rrk = reinterpret_cast<u64 *>(dr_smem + a.o_rrk);    // rerank keys alias a region that is dead after the traversal
    const bool no_smem_hash = RW8 >= 3 ? (DR_L2V != 0) : (a.hash_cap == 0);

    __shared__ long long s_b;
    __shared__ u64 s_pfkey;   // prefetch == 2: a survivor below this key is among the next step's likely expansions
    __shared__ __align__(8) uint64_t s_lutbar;
    __shared__ __align__(8) uint64_t s_rrbar[16];   // rerank staging: two half-row barriers per warp
    __shared__ __align__(8) uint64_t s_adjbar[16];  // prefetch & 16: one barrier per selected node's adjacency row
    __shared__ int s_nn2[2], s_ns, s_nspec, s_mvalid, s_hcount, s_ovfcount, s_ovfused, s_status;

    const int tid = threadIdx.x;
    constexpr int nt = DR_FAST_NT, nw = DR_FAST_NT / 32;   // the launcher always uses DR_FAST_NT threads
    int lane = tid & 31, wid = tid >> 5;
    asm volatile("" : "+r"(lane), "+r"(wid));    // opaque: keep them in registers instead of re-reading %tid in the loops
    const uint32_t lt_mask = (1u << lane) - 1u;
    const uint32_t tab32 = smem_u32(s_lut);
    // this lane's word of code row 0 (full chunk / tail chunk): kept in registers, see rows_load
    const uint8_t *code_lane = a.codes + lane * 4;
    const uint8_t *code_tail = a.codes + 128 * (((WORDS > 0 ? WORDS : (a.M >> 2))) >> 5) + (lane & 15) * 4;
    asm volatile("" : "+l"(code_lane), "+l"(code_tail));
    const int D = RW8 >= 2 ? 1536 : a.D, M = a.M, L = RW8 >= 3 ? 100 : a.L;
    const uint32_t hcap = RW8 >= 3 ? (DR_L2V ? 0u : 4096u) : a.hash_cap;
    const int pf = RW8 == 4 ? DR_PF_SPEC : a.prefetch;
    const bool adj_async = (pf & 16) != 0;   // the launcher clears the bit unless R % 4 == 0 and W <= 16
    const bool spec_code = (pf & 8) != 0;
    uint32_t adj_par = 0u;                   // bit s: parity of s_adjbar[s] this warp waits for next
    // Selection record: the merge writes the list positions of the first DR_SELCAP unexpanded entries, by rank.  The next steps take
    // their entries from it by rank offset, so a step that follows a step WITHOUT survivors (list unchanged) needs neither a rescan of
    // the list nor a block barrier: every thread derives the step's size from ur[n] and its own count of expansions.
    const bool selrec = (pf & (8 | 16)) == 0;
    const uint8_t *deleted = RW8 == 4 ? nullptr : a.deleted;
    const bool do_rerank = RW8 == 4 ? true : (a.rerank != 0);
    const int R = RW8 ? 32 : a.R, W = RW8 ? 8 : a.W;
    // A step none of whose newcomers entered the list leaves the list as it was, so the entries the NEXT steps expand are already
    // known: the following step expands up to W2 of them at once (same expansions, same claims, fewer barrier-separated steps)
    const int W2 = RW8 ? DR_W2_SPEC : a.W2;
    const int words = WORDS > 0 ? WORDS : (M >> 2);
    const uint32_t hmask = hcap ? hcap - 1u : 0u, ovf_mask = a.ovf_cap - 1u;
    const uint32_t hshift = 32u - (uint32_t)__popc(hmask), ovf_shift = 32u - (uint32_t)__popc(ovf_mask);
    const int hlimit = hcap ? (int)(hcap - (hcap >> 2)) : -1;   // no table: "full" from the start
    const int ovf_limit = (int)(a.ovf_cap - (a.ovf_cap >> 2));
    uint32_t *my_ovf = a.ovf + (size_t)blockIdx.x * a.ovf_cap;
    const uint32_t n32 = a.N > 0xFFFFFFFFll ? 0xFFFFFFFFu : (uint32_t)a.N;   // ids are 32-bit; DR_EMPTY is never valid
    uint32_t lut_phase = 0;
#ifdef DR_PHASE_TIMING
    long long pt_acc[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, pt_last = clock64();
#endif
    if (tid == 0) {
        mbar_init(&s_lutbar, 1);
        for (int i = 0; i < 16; ++i) { mbar_init(&s_rrbar[i], 1); mbar_init(&s_adjbar[i], 1); }
        fence_mbar_init();
    }
    uint32_t rr_ph0 = 0, rr_ph1 = 0;
    const uint64_t pol_stream = l2_policy_evict_first();

    if (tid == 0) s_b = (long long)atomicAdd(a.counter, 1ull);
    for (;;) {
        __syncthreads();
        DR_PT(0);   // the previous query's output
        const long long b = s_b;      // fetched during the previous query (its table is already on its way to L2)
        if (b >= a.B) break;

        // ---- stage the query's table (permuting load: global [w][c][4] -> bank-per-lane layout), the query, the hash
        if (tid == 0) { s_hcount = 1; s_ovfcount = no_smem_hash ? 1 : 0; s_ovfused = no_smem_hash ? 1 : 0; s_status = 0; s_nn2[0] = 0; }
        if (WP) {
            const int nfull = words >> 5, rem = words & 31;
            const uint4 *src = reinterpret_cast<const uint4 *>(a.lut8 + (size_t)b * M * 256);
            const int total = words * 32;          // unit = 8 centroids of one code word (2 x 16 B)
            constexpr int UB = 3;
            for (int u0 = tid; u0 < total; u0 += nt * UB) {
                uint4 v[UB][2];
#pragma unroll
                for (int i = 0; i < UB; ++i) {
                    const int u = u0 + i * nt;
                    if (u < total) {
                        const int cg = u / words, w = u - cg * words;   // consecutive lanes -> consecutive words (banks)
                        const uint4 *p = src + (w * 64 + cg * 2);
                        v[i][0] = __ldg(p); v[i][1] = __ldg(p + 1);
                    }
                }
#pragma unroll
                for (int i = 0; i < UB; ++i) {
                    const int u = u0 + i * nt;
                    if (u < total) {
                        const int cg = u / words, w = u - cg * words;
                        uint32_t *dst; int sw;
                        if (w < nfull * 32) { dst = reinterpret_cast<uint32_t *>(s_lut + (w >> 5) * 32768) + (w & 31); sw = 32; }
                        else { dst = reinterpret_cast<uint32_t *>(s_lut + nfull * 32768) + (w - nfull * 32); sw = rem; }
                        dst += cg * 8 * sw;
                        dst[0] = v[i][0].x; dst[sw] = v[i][0].y; dst[2 * sw] = v[i][0].z; dst[3 * sw] = v[i][0].w;
                        dst[4 * sw] = v[i][1].x; dst[5 * sw] = v[i][1].y; dst[6 * sw] = v[i][1].z; dst[7 * sw] = v[i][1].w;
                    }
                }
            }
        } else if (tid == 0) {
            fence_proxy_async();   // the region was last touched through the generic proxy
            mbar_expect_tx(&s_lutbar, (uint32_t)M * 256u);
            bulk_g2s(s_lut, a.lut8 + (size_t)b * M * 256, (uint32_t)M * 256u, &s_lutbar);
        }
        for (uint32_t i = tid; i < hcap; i += nt) s_hash[i] = DR_EMPTY;
        if (!WP) { mbar_wait(&s_lutbar, lut_phase); lut_phase ^= 1u; }
        __syncthreads();

        if (wid == 0) {
            uint32_t s0;
            if (WP) {
                const uint32_t gid[2] = {a.start, a.start};
                uint32_t gs[2];
                adc_u8_rows<(WORDS > 0 ? WORDS : 0), 2>(a.codes, M, s_lut, gid, lane, gs);
                s0 = gs[0];
            } else {
                s0 = adc_u8_warp(a.codes + (size_t)a.start * M, s_lut, M, lane);
            }
            if (lane == 0) {
                s_list0[0] = make_ikey(s0, a.start);
                if (no_smem_hash) my_ovf[fib_slot(a.start ^ 0x9e3779b9u, ovf_shift)] = a.start;
                else s_hash[fib_slot(a.start, hshift)] = a.start;
                s_ur0[0] = 0; s_ur0[1] = 1;      // one unexpanded entry
                s_sel[selrec ? 0 : W] = 0u; s_ns = 1;         // the first step expands list position 0
                s_nspec = 0;
                s_pfkey = DR_KEY_MAX;
                if (adj_async) {
                    mbar_expect_tx(&s_adjbar[0], (uint32_t)R * 4u);
                    bulk_g2s(s_adjrow, a.adj + (size_t)a.start * R, (uint32_t)R * 4u, &s_adjbar[0]);
                }
            }
        }
        int cur = 0, n = 1, hops = 0, nvis = 1, step = 0;
        int ubase = 0;    // entries marked expanded since the ur[] array of the current list was written
        int sel_base = 0; // selrec: ubase at the time s_sel[0] was (re)written
        int ns_reg = 1;   // entries the coming step expands
        bool mv_dirty = true;   // s_mvalid has to be reset before this step's ADC phase counts survivors
        __syncthreads();
        DR_PT(1);   // table / query staging, start node

        for (;;) {
            u64 *lst = cur ? s_list1 : s_list0;
            u64 *oth = cur ? s_list0 : s_list1;
            // (1+2) The selection (list positions of the first W unexpanded entries, s_sel / s_ns) was produced by the
            //       previous step's merge: warp s loads that node's adjacency row and claims its first-seen neighbours.
            uint16_t *ur_old = cur ? s_ur1 : s_ur0;
            uint16_t *ur_new = cur ? s_ur0 : s_ur1;
            int *p_nn = &s_nn2[step & 1];
            const int ns = ns_reg;
            if (ns == 0) break;
            const bool use_ovf_now = (s_hcount + W2 * R > hlimit);   // same value for every thread (read after the last barrier)
            if (use_ovf_now && (s_ovfcount + W2 * R > ovf_limit)) {
                if (tid == 0) s_status |= DR_ST_VISITED_OVERFLOW;
                if (adj_async)      // rows already on their way: consume their barrier phases so the next query starts in step
                    for (int s = wid; s < ns; s += nw) { mbar_wait(&s_adjbar[s], (adj_par >> s) & 1u); adj_par ^= 1u << s; }
                break;
            }
            if (tid == 0) {
                if (mv_dirty) s_mvalid = 0;   // (after a barrier-free empty step it is still 0, and a slower warp may not have read it yet)
                if (use_ovf_now) s_ovfused = 1;
            }
            const bool spec = (pf & 2) != 0;
            // prefetch & 8: the nodes next in line (unexpanded ranks W .. 2W-1) are this step's likely successors: read their
            // adjacency rows now (L2 hits: prefetched when they were accepted) and start the code rows of their unseen
            // neighbours on the trip to L2, one whole step before the ADC phase that needs them.  Nothing is claimed.
            const int nspec = (spec_code && !use_ovf_now && !no_smem_hash) ? s_nspec : 0;
            uint32_t nb2 = DR_EMPTY;
            if (wid < nspec && lane < R) nb2 = __ldg(a.adj + (size_t)key_id(lst[s_sel[wid]]) * R + lane);
            for (int s = wid; s < ns; s += nw) {
                const int pos = (int)s_sel[selrec ? (ubase - sel_base + s) : (W + s)];
                DR_PT(6);   // (timing build) selection read
                const uint32_t node = key_id(lst[pos]);
                __syncwarp();
                if (lane == 0) lst[pos] |= 1ull;          // mark it expanded: only this warp touches lst[pos] before the barrier
                const uint32_t *row = a.adj + (size_t)node * R;
                if (adj_async) {                          // the row was sent to shared memory when the node was selected
                    mbar_wait(&s_adjbar[s], (adj_par >> s) & 1u);
                    adj_par ^= 1u << s;
                }
                for (int j0 = 0; j0 < R; j0 += 32) {
                    const int j = j0 + lane;
                    const uint32_t nb = (j < R) ? (adj_async ? s_adjrow[s * R + j] : __ldg(row + j)) : DR_EMPTY;
                    bool valid = (j < R) && (nb < n32);
#ifdef DR_PHASE_TIMING
                    asm volatile("" ::"r"(nb));
                    DR_PT(7);   // (timing build) adjacency row arrived
#endif
                    if (valid && deleted) valid = deleted[nb] == 0;
                    bool isnew = false;   // equal ids in one row (0-padding): the CAS admits exactly one of them
                    if (valid) isnew = visited_insert_fast(nb, s_hash, hmask, hshift, use_ovf_now, my_ovf, ovf_mask, ovf_shift, no_smem_hash);
                    const unsigned m = __ballot_sync(DR_FULL, isnew);
                    const int cnt = __popc(m);
                    int basepos = 0;
                    if (cnt) {
                        if (lane == 0) basepos = atomicAdd(p_nn, cnt);
                        basepos = __shfl_sync(DR_FULL, basepos, 0);
                    }
                    if (isnew) {
                        s_newid[basepos + __popc(m & lt_mask)] = nb;
                        if (pf & 4) {   // the code row is needed right after the barrier: start its trip now
                            const uint8_t *cr = a.codes + (size_t)nb * M;
                            asm volatile("prefetch.global.L2 [%0];" ::"l"(cr));
                            asm volatile("prefetch.global.L2 [%0];" ::"l"(cr + (M > 128 ? 128 : M - 1)));
                        }
                    }
                }
            }
            for (int s2 = wid; s2 < nspec; s2 += nw) {
                const uint32_t *row2 = a.adj + (size_t)key_id(lst[s_sel[s2]]) * R;
                for (int j0 = 0; j0 < R; j0 += 32) {
                    const int j = j0 + lane;
                    const uint32_t nb = (s2 == wid && j0 == 0) ? nb2 : ((j < R) ? __ldg(row2 + j) : DR_EMPTY);
                    if (j < R && nb < n32) {
                        bool seen = false;                 // read-only probe (a concurrent claim may be missed: it is only a hint)
                        for (uint32_t h = fib_slot(nb, hshift);; h = (h + 1) & hmask) {
                            const uint32_t c = s_hash[h];
                            if (c == nb) { seen = true; break; }
                            if (c == DR_EMPTY) break;
                        }
                        if (!seen) {
                            const uint8_t *cr = a.codes + (size_t)nb * M;
                            asm volatile("prefetch.global.L2 [%0];" ::"l"(cr));
                            asm volatile("prefetch.global.L2 [%0];" ::"l"(cr + (M > 128 ? 128 : M - 1)));
                        }
                    }
                }
            }
            __syncthreads();
            DR_PT(2);   // select + adjacency + visited
            const bool use_ovf = use_ovf_now;
            ubase += ns;
            if (tid == 0) s_nn2[(step + 1) & 1] = 0;          // next step's newcomer counter
            ++step;
            // (3) quantised ADC of the newcomers; the survivors are appended compactly
            const int nn = *p_nn;
            const bool full = (n >= L);
            const u64 worstk = lst[n - 1] & ~1ull;
            const u64 pfkey = s_pfkey;
            if (WP) {
                // item i of this warp = newcomer wid + nw * i; lane l owns item l of the current block of 32 items, rows go
                // through the table four (or two) at a time, and the warp appends its survivors once per block
                constexpr int KW = (WORDS > 0 ? WORDS : 0);
#ifdef DR_P2_CONTIG
                for (int base = 0; base < nn; base += 32 * nw) {     // warp w takes a contiguous run of the block's items
                    const int cnt_blk = (nn - base) < 32 * nw ? (nn - base) : 32 * nw;
                    const int per = (cnt_blk + nw - 1) / nw;
                    int cntw = cnt_blk - wid * per;
                    cntw = cntw < per ? cntw : per;
                    if (cntw <= 0) continue;
                    const uint32_t myid = lane < cntw ? s_newid[base + wid * per + lane] : 0u;
#else
                for (int first = wid; first < nn; first += 32 * nw) {
                    int cntw = (nn - first + nw - 1) / nw;
                    cntw = cntw < 32 ? cntw : 32;
                    const uint32_t myid = lane < cntw ? s_newid[first + lane * nw] : 0u;
#endif
                    uint32_t mysum = 0u;
                    const int rounds = (cntw + 3) >> 2;
                    RowWords<4> wa, wb;   // two groups of code words: one being summed, one in flight
                    {
                        uint32_t gid[4];
#pragma unroll
                        for (int g = 0; g < 4; ++g) gid[g] = __shfl_sync(DR_FULL, myid, g);    // lanes >= cntw hold row 0: valid
                        rows_load<KW, 4>(code_lane, code_tail, M, gid, lane, wa);
                    }
                    for (int r = 0; r < rounds; r += 2) {
                        if (r + 1 < rounds) {
                            uint32_t gid[4];
#pragma unroll
                            for (int g = 0; g < 4; ++g) gid[g] = __shfl_sync(DR_FULL, myid, (r + 1) * 4 + g);
                            rows_load<KW, 4>(code_lane, code_tail, M, gid, lane, wb);
                        }
                        uint32_t gs[4];
                        rows_sum<KW, 4>(M, tab32, wa, (cntw - r * 4) > 2 ? 2 : 1, lane, gs);
#ifdef DR_PHASE_TIMING
                        if (r == 0) { asm volatile("" ::"r"(gs[0])); DR_PT(8); }   // (timing build) first group of code rows summed
#endif
                        {   // lane 4 r + g keeps row g of this round (two selects on loop-invariant predicates, one on r)
                            const uint32_t lo2 = (lane & 1) ? gs[1] : gs[0], hi2 = (lane & 1) ? gs[3] : gs[2];
                            const uint32_t pick = (lane & 2) ? hi2 : lo2;
                            mysum = ((lane >> 2) == r) ? pick : mysum;
                        }
                        if (r + 1 < rounds) {
                            if (r + 2 < rounds) {
                                uint32_t gid[4];
#pragma unroll
                                for (int g = 0; g < 4; ++g) gid[g] = __shfl_sync(DR_FULL, myid, (r + 2) * 4 + g);
                                rows_load<KW, 4>(code_lane, code_tail, M, gid, lane, wa);
                            }
                            rows_sum<KW, 4>(M, tab32, wb, (cntw - (r + 1) * 4) > 2 ? 2 : 1, lane, gs);
                            {
                                const uint32_t lo2 = (lane & 1) ? gs[1] : gs[0], hi2 = (lane & 1) ? gs[3] : gs[2];
                                const uint32_t pick = (lane & 2) ? hi2 : lo2;
                                mysum = ((lane >> 2) == r + 1) ? pick : mysum;
                            }
                        }
                    }
                    const u64 key = make_ikey(mysum, myid);
                    const bool ok = lane < cntw && (!full || key < worstk);
                    const unsigned okm = __ballot_sync(DR_FULL, ok);
                    if (okm) {
                        int basep = 0;
                        if (lane == 0) basep = atomicAdd(&s_mvalid, __popc(okm));
                        basep = __shfl_sync(DR_FULL, basep, 0);
                        if (ok) {
                            s_newk[basep + __popc(okm & lt_mask)] = key;
                            if ((pf & 1) || ((pf & 2) && key < pfkey))   // likely to be expanded next step
                                for (int o = 0; o < R; o += 32) asm volatile("prefetch.global.L2 [%0];" ::"l"(a.adj + (size_t)myid * R + o));
                        }
                    }
                }
            } else {
                for (int i = wid; i < nn; i += nw) {
                    const uint32_t id = s_newid[i];
                    const uint32_t sum = adc_u8_warp(a.codes + (size_t)id * M, s_lut, M, lane);
                    if (lane == 0) {
                        const u64 key = make_ikey(sum, id);
                        if (!full || key < worstk) s_newk[atomicAdd(&s_mvalid, 1)] = key;
                    }
                }
            }
            if (tid == 0) {
                if (use_ovf) s_ovfcount += nn; else s_hcount += nn;
            }
            __syncthreads();
            DR_PT(3);   // ADC
            nvis += nn;
            hops += ns;
            const int mv = s_mvalid;
            // (4) merge (keys are unique, so ranks are exact):
            //     few survivors   : rank = count of smaller newcomers (no sort, no extra barrier)
            //     up to 64 per warp: every warp bitonic-sorts one 64-key chunk in registers, then every item sums its
            //                        binary-search ranks in the other sorted sequences
            //     more            : rank counting over all newcomers
            //     Every placed element also gets its unexpanded rank (unexpanded old entries before it, from ur_old, plus
            //     the newcomers before it): ranks < W are the next step's selection, so no step ever scans the list.
            const int total = n + mv;
            const int nnew = total < L ? total : L;
            auto place = [&](u64 key, int pos, int from, bool isnew) {
                if (pos >= L) return;
                const int ub = (int)ur_old[from] - ubase;
                const int rank = (ub > 0 ? ub : 0) + (pos - from);
                const bool un = isnew || !(key & 1ull);
                oth[pos] = key;
                ur_new[pos] = (uint16_t)rank;
                if (un) {
                    if (selrec) {
                        if (rank < DR_SELCAP) s_sel[rank] = (uint32_t)pos;
                    } else if (rank < W) {
                        s_sel[W + rank] = (uint32_t)pos;
                        if (adj_async) {                 // its adjacency row travels to shared memory under the rest of the merge
                            mbar_expect_tx(&s_adjbar[rank], (uint32_t)R * 4u);
                            bulk_g2s(s_adjrow + rank * R, a.adj + (size_t)key_id(key) * R, (uint32_t)R * 4u, &s_adjbar[rank]);
                        }
                    } else if (rank < 2 * W) {           // next in line: likely to be expanded in two steps
                        if (spec_code) s_sel[rank - W] = (uint32_t)pos;
                        if (spec) {
                            for (int o = 0; o < R; o += 32) asm volatile("prefetch.global.L2 [%0];" ::"l"(a.adj + (size_t)key_id(key) * R + o));
                            if (rank == 2 * W - 1) s_pfkey = key;
                        }
                    }
                }
                if (pos == nnew - 1) {
                    const int tot = rank + (un ? 1 : 0);
                    ur_new[nnew] = (uint16_t)tot;
                    s_ns = tot < W ? tot : W;
                    s_nspec = tot < W ? 0 : (tot < 2 * W ? tot - W : W);
                    if (spec && tot < 2 * W) s_pfkey = DR_KEY_MAX;
                }
            };
            if (mv > 0) {
                if (mv <= DR_MERGE_LINEAR || mv > 64 * nw) {
#if DR_MERGE_PAIRS
                    // two adjacent lanes share an item: each counts the smaller newcomers over half of them, so that a step with
                    // ~100 + mv items keeps all eight warps busy and the scan every item waits for is half as long
                    const int hsel = tid & 1, mh = mv >> 1;
                    const int j0 = hsel ? mh : 0, j1 = hsel ? mv : mh;
                    for (int x0 = 0; x0 < total; x0 += nt / 2) {
                        const int x = x0 + (tid >> 1);
                        u64 key = 0ull;
                        int from = 0, cnt = 0;
                        if (x < total) {
                            if (x < n) { key = lst[x]; from = x; }
                            else { key = s_newk[x - n]; from = lower_bound_u64<(RW8 >= 3 ? 64 : 512)>(lst, n, key); }
                            for (int j = j0; j < j1; ++j) cnt += (s_newk[j] < key) ? 1 : 0;
                        }
                        cnt += __shfl_xor_sync(DR_FULL, cnt, 1);
                        if (x < total && hsel == 0) place(key, from + cnt, from, x >= n);
                    }
#else
                    for (int x = tid; x < total; x += nt) {
                        u64 key;
                        int pos, from;
                        if (x < n) { key = lst[x]; from = x; }
                        else { key = s_newk[x - n]; from = lower_bound_u64<(RW8 >= 3 ? 64 : 512)>(lst, n, key); }
                        pos = from;
                        for (int j = 0; j < mv; ++j) pos += (s_newk[j] < key) ? 1 : 0;
                        place(key, pos, from, x >= n);
                    }
#endif
                } else {
                    // every warp bitonic-sorts one 64-key chunk in registers (two keys per lane: element e = lane + 32 r),
                    // then every item sums its binary-search ranks in the other sorted sequences
                    const int nch = (mv + 63) >> 6;
                    if (wid < nch) {
                        const int i0 = (wid << 6) + lane, i1 = i0 + 32;
                        u64 k0 = i0 < mv ? s_newk[i0] : DR_KEY_MAX, k1 = i1 < mv ? s_newk[i1] : DR_KEY_MAX;
#pragma unroll
                        for (int k2 = 2; k2 <= 64; k2 <<= 1) {
#pragma unroll
                            for (int j = k2 >> 1; j > 0; j >>= 1) {
                                if (j == 32) {                                   // partners sit in the same lane; k2 == 64: ascending
                                    const u64 mn = k0 < k1 ? k0 : k1, mx = k0 < k1 ? k1 : k0;
                                    k0 = mn; k1 = mx;
                                } else {
                                    const u64 o0 = __shfl_xor_sync(DR_FULL, k0, j), o1 = __shfl_xor_sync(DR_FULL, k1, j);
                                    const bool lower = ((lane & j) == 0);          // this element is the smaller index of its pair
                                    const bool up0 = k2 == 64 ? true : ((lane & k2) == 0);                       // e = lane
                                    const bool up1 = k2 == 64 ? true : (k2 == 32 ? false : ((lane & k2) == 0));    // e = lane + 32
                                    k0 = (up0 == lower) ? (k0 < o0 ? k0 : o0) : (k0 < o0 ? o0 : k0);
                                    k1 = (up1 == lower) ? (k1 < o1 ? k1 : o1) : (k1 < o1 ? o1 : k1);
                                }
                            }
                        }
                        if (i0 < mv) s_newk[i0] = k0;
                        if (i1 < mv) s_newk[i1] = k1;
                    }
                    __syncthreads();
                    for (int x = tid; x < total; x += nt) {
                        u64 key;
                        int pos, from, own = -1;
                        if (x < n) { key = lst[x]; from = x; pos = x; }
                        else {
                            const int j = x - n;
                            key = s_newk[j];
                            own = j >> 6;
                            from = lower_bound_u64<(RW8 >= 3 ? 64 : 512)>(lst, n, key);
                            pos = (j & 63) + from;
                        }
                        for (int c = 0; c < nch; ++c) {
                            if (c == own) continue;
                            const int sz = (mv - (c << 6)) < 64 ? (mv - (c << 6)) : 64;
                            pos += lower_bound_u64<64>(s_newk + (c << 6), sz, key);
                        }
                        place(key, pos, from, x >= n);
                    }
                }
                cur ^= 1;
                n = nnew;
                ubase = 0;
                sel_base = 0;
                __syncthreads();   // the next step reads the merged list and its selection
                ns_reg = s_ns;
                mv_dirty = true;
                DR_PT(4);   // merge
            } else {
                // nothing survived: the list stays, the selection moves on by the ns entries just expanded
                int tot_left = (int)ur_old[n] - ubase;
                tot_left = tot_left > 0 ? tot_left : 0;
                const int ns_next = tot_left < W2 ? tot_left : W2;
                if (selrec && ubase - sel_base + ns_next <= DR_SELCAP) {
                    ns_reg = ns_next;          // the recorded ranks cover the step: no rescan, no barrier
                    mv_dirty = false;          // nothing survived: s_mvalid is 0 and stays untouched until the next ADC phase
                    continue;
                }
                for (int x = tid; x < n; x += nt) {
                    const int r = (int)ur_old[x] - ubase;
                    if (!(lst[x] & 1ull)) {
                        if (selrec) {
                            if (r < DR_SELCAP) s_sel[r] = (uint32_t)x;
                        } else if (r < W2) {
                            s_sel[W + r] = (uint32_t)x;
                            if (adj_async) {
                                mbar_expect_tx(&s_adjbar[r], (uint32_t)R * 4u);
                                bulk_g2s(s_adjrow + r * R, a.adj + (size_t)key_id(lst[x]) * R, (uint32_t)R * 4u, &s_adjbar[r]);
                            }
                        } else if (W2 == W && r < 2 * W) {
                            if (spec_code) s_sel[r - W] = (uint32_t)x;
                            if (spec) {
                                for (int o = 0; o < R; o += 32) asm volatile("prefetch.global.L2 [%0];" ::"l"(a.adj + (size_t)key_id(lst[x]) * R + o));
                                if (r == 2 * W - 1) s_pfkey = lst[x];
                            }
                        }
                    }
                }
                if (tid == 0) {
                    int tot = (int)ur_old[n] - ubase;
                    tot = tot > 0 ? tot : 0;
                    s_ns = tot < W2 ? tot : W2;
                    s_nspec = (W2 != W || tot < W) ? 0 : (tot < 2 * W ? tot - W : W);
                    if (spec && tot < 2 * W) s_pfkey = DR_KEY_MAX;
                }
                sel_base = ubase;
                __syncthreads();
                ns_reg = s_ns;
                mv_dirty = true;
                DR_PT(4);
            }
        }

        const u64 *lst = cur ? s_list1 : s_list0;
        const float sc = a.lut_scale[b], off = a.lut_offset[b];
        if (tid == 0) s_b = (long long)atomicAdd(a.counter, 1ull);   // next query (every thread holds b in a register by now)
        if (a.list_ids) {
            for (int i = tid; i < L; i += nt) {
                a.list_ids[(size_t)b * L + i] = i < n ? (int32_t)key_id(lst[i]) : -1;
                if (a.list_dist)
                    a.list_dist[(size_t)b * L + i] = i < n ? __fmaf_rn(sc, (float)(uint32_t)(lst[i] >> 32), off) : __int_as_float(0x7f800000);
            }
        }
        const int k = a.k;
        u64 *peer_dst = nullptr;      // this query's k slots in the owner rank's receive buffer (peer memory)
        if (a.peer_recv) {
            int g; long long r;
            peer_slot(a.peer_q0 + b, a.peer_B, a.peer_G, g, r);
            const long long Bq = (a.peer_B + a.peer_G - 1) / a.peer_G;
            peer_dst = a.peer_recv[g] + ((size_t)a.peer_rank * Bq + r) * k;
        }
        if (do_rerank) {
            // Exact rerank of the whole list.  The table is dead now: its bytes stage full-precision rows, one slot per
            // warp, filled by bulk async copies (TMA engine, evict-first in L2) in two pieces so that the second half of a
            // row and the first half of the next one are in flight while the warp accumulates; no registers are tied up
            // by loads in flight and 8 rows x 6 KB per CTA keep HBM busy.  Same arithmetic order as warp_l2sq.
            const int slot = (D * 4 + 15) & ~15;
            int nsl = a.rr_slots;
            nsl = nsl < nw ? nsl : nw;
            const bool staged = nsl >= 1 && (D & 3) == 0;
            float *buf = reinterpret_cast<float *>(s_lut + wid * slot);
            uint64_t *bar0 = &s_rrbar[2 * (wid & 7)], *bar1 = bar0 + 1;
            const int e0 = (((D + 127) >> 7) >> 1) << 7;     // elements in the first piece (multiple of 128, may be 0)
            const uint32_t by0 = (uint32_t)e0 * 4u, by1 = (uint32_t)(D - e0) * 4u;
            if (staged && wid < nsl && lane == 0 && wid < n) {   // first rows in flight before anything else
                const float *row = a.vec + (size_t)key_id(lst[wid]) * D;
                fence_proxy_async();
                if (by0) { mbar_expect_tx(bar0, by0); bulk_g2s_hint(buf, row, by0, bar0, pol_stream); }
                mbar_expect_tx(bar1, by1); bulk_g2s_hint(buf + e0, row + e0, by1, bar1, pol_stream);
            }
#if DR_RR_PF
            if (staged) {
                const int pf_end = DR_RR_PF == 1 ? n : (n < (1 + DR_RR_PF) * nsl ? n : (1 + DR_RR_PF) * nsl);
                for (int i = nsl + tid; i < pf_end; i += nt) bulk_prefetch_l2(a.vec + (size_t)key_id(lst[i]) * D, (uint32_t)D * 4u);
            }
#endif
            {   // the query vector goes where the (dead) visited table was
                const float *qg = a.Q + (size_t)b * D;
                for (int i = tid; i < D; i += nt) s_q[i] = __ldg(qg + i);
            }
            __syncthreads();
            {   // start the next query's table on its trip to L2: the staging loop of the next iteration hits there
                const long long bn = s_b;
                if (bn < a.B) {
                    const uint8_t *tn = a.lut8 + (size_t)bn * M * 256;
                    for (int o = tid * 128; o < M * 256; o += nt * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(tn + o));
                }
            }
            if (staged) {
                if (wid < nsl) {
                    for (int i = wid; i < n; i += nsl) {
                        const int inext = i + nsl;
                        const float *rown = inext < n ? a.vec + (size_t)key_id(lst[inext]) * D : nullptr;
#if DR_RR_PF >= 2
                        if (lane == 0 && i + (1 + DR_RR_PF) * nsl < n)
                            bulk_prefetch_l2(a.vec + (size_t)key_id(lst[i + (1 + DR_RR_PF) * nsl]) * D, (uint32_t)D * 4u);
#endif
                        float acc = 0.0f;
                        if (by0) {
                            mbar_wait(bar0, rr_ph0); rr_ph0 ^= 1u;
                            acc = l2sq_piece_smem<(RW8 >= 2 ? DR_RR_FULL : 0)>(buf, s_q, 0, e0, lane, acc);
                            __syncwarp();
                            if (lane == 0 && rown) { mbar_expect_tx(bar0, by0); bulk_g2s_hint(buf, rown, by0, bar0, pol_stream); }
                        }
                        mbar_wait(bar1, rr_ph1); rr_ph1 ^= 1u;
                        acc = l2sq_piece_smem<(RW8 >= 2 ? DR_RR_FULL : 0)>(buf, s_q, e0, D, lane, acc);
                        __syncwarp();
                        if (lane == 0 && rown) { mbar_expect_tx(bar1, by1); bulk_g2s_hint(buf + e0, rown + e0, by1, bar1, pol_stream); }
                        const float d2 = warp_sum_butterfly(acc);
                        if (lane == 0) s_rrk[i] = ((u64)f2ord(d2 + 0.0f) << 32) | (u64)i;
                    }
                }
            } else {
                for (int i = wid; i < n; i += 2 * nw) {
                    const int i2 = i + nw;
                    if (i2 < n) {
                        float dA, dB;
                        warp_l2sq_x2(a.vec + (size_t)key_id(lst[i]) * D, a.vec + (size_t)key_id(lst[i2]) * D, s_q, D, lane, dA, dB);
                        if (lane == 0) {
                            s_rrk[i] = ((u64)f2ord(dA + 0.0f) << 32) | (u64)i;
                            s_rrk[i2] = ((u64)f2ord(dB + 0.0f) << 32) | (u64)i2;
                        }
                    } else {
                        float d2 = warp_l2sq(a.vec + (size_t)key_id(lst[i]) * D, s_q, D, lane);
                        if (lane == 0) s_rrk[i] = ((u64)f2ord(d2 + 0.0f) << 32) | (u64)i;
                    }
                }
            }
            __syncthreads();
            DR_PT(5);   // rerank distance pass
            // rank of every reranked entry among the n (keys are unique: (distance bits, list position)); with n <= 128 and 256
            // threads two threads share an entry, each counting over half of the keys, so that all eight warps work on the pass
            const bool split = (nt == 256) && n <= 128;
            uint32_t *s_part = s_newid;                      // dead after the traversal: the upper halves' partial ranks
            for (int i0 = 0; i0 < n; i0 += (split ? 128 : nt)) {
                const int i = i0 + (split ? (tid & 127) : tid);
                const int hi = split ? (tid >> 7) : 0;
                const int j0 = (split && hi) ? (n >> 1) : 0, j1 = (split && !hi) ? (n >> 1) : n;
                u64 key = 0ull;
                int pos = 0;
                if (i < n) {
                    key = s_rrk[i];
                    for (int j = j0; j < j1; ++j) pos += (s_rrk[j] < key) ? 1 : 0;
                    if (split && hi) s_part[i] = (uint32_t)pos;
                }
                if (split) __syncthreads();
                if (i < n && !hi) {
                    if (split) pos += (int)s_part[i];
                    if (pos < k) {
                        float d2 = ord2f((uint32_t)(key >> 32));
                        a.out_ids[(size_t)b * k + pos] = (int32_t)key_id(lst[i]);
                        if (a.out_dist) a.out_dist[(size_t)b * k + pos] = a.sqrt_out ? sqrtf(d2) : d2;
                        if (peer_dst) peer_dst[pos] = (key & 0xFFFFFFFF00000000ull) | (u64)(uint32_t)(key_id(lst[i]) + a.peer_id_offset);
                    }
                }
            }
        } else {
            for (int i = tid; i < n && i < k; i += nt) {
                float d = __fmaf_rn(sc, (float)(uint32_t)(lst[i] >> 32), off);
                a.out_ids[(size_t)b * k + i] = (int32_t)key_id(lst[i]);
                if (a.out_dist) a.out_dist[(size_t)b * k + i] = a.sqrt_out ? sqrtf(fmaxf(d, 0.0f)) : d;
            }
        }
        for (int i = n + tid; i < k; i += nt) {
            a.out_ids[(size_t)b * k + i] = -1;
            if (a.out_dist) a.out_dist[(size_t)b * k + i] = __int_as_float(0x7f800000);
            if (peer_dst) peer_dst[i] = DR_KEY_MAX;
        }
        if (tid == 0) {
            if (a.out_hops) a.out_hops[b] = hops;
            if (a.out_visited) a.out_visited[b] = nvis;
            if (a.list_len) a.list_len[b] = n;
            if (a.status) a.status[b] = s_status;
        }
        if (s_ovfused) {
            for (uint32_t i = tid; i < a.ovf_cap; i += nt) my_ovf[i] = DR_EMPTY;
        }
    }
#ifdef DR_PHASE_TIMING
    if (tid == 0)
        for (int i = 0; i < 12; ++i) atomicAdd(a.counter + 1 + i, (u64)pt_acc[i]);
#endif
}

typedef void (*fast_kernel_t)(const FastArgs);

template <int MINB, int RW8>
static fast_kernel_t pick_fast_kernel_b(int M) {
    if ((M & 3) != 0 || M > 256) return search_fast_kernel<-1, MINB, RW8>;
    switch (M >> 2) {
        case 16: return search_fast_kernel<16, MINB, RW8>;   // M = 64  (the adaptive default at D = 1536, adaptive_pq.py:81-108)
        case 32: return search_fast_kernel<32, MINB, RW8>;   // M = 128
        case 48: return search_fast_kernel<48, MINB, RW8>;   // M = 192
        case 64: return search_fast_kernel<64, MINB, RW8>;   // M = 256
        default: return search_fast_kernel<0, MINB, RW8>;
    }
}
static fast_kernel_t pick_fast_kernel(int M, int minb, int rw8) {
#if DR_L2V
    if (rw8 == 4 && minb >= 4) return pick_fast_kernel_b<4, 4>(M);
    if (rw8 >= 3 && minb >= 4) return pick_fast_kernel_b<4, 3>(M);
#else
    if (rw8 == 4 && minb < 4) return pick_fast_kernel_b<3, 4>(M);
    if (rw8 >= 3 && minb < 4) return pick_fast_kernel_b<3, 3>(M);
#endif
    if (rw8 >= 2) return minb >= 4 ? pick_fast_kernel_b<4, 2>(M) : pick_fast_kernel_b<3, 2>(M);
    if (rw8) return minb >= 4 ? pick_fast_kernel_b<4, 1>(M) : pick_fast_kernel_b<3, 1>(M);
    return minb >= 4 ? pick_fast_kernel_b<4, 0>(M) : pick_fast_kernel_b<3, 0>(M);
}

int launch_search_fast(dr_index *h, const float *d_Q, int64_t B, const dr_search_params *p, int32_t *ids, float *dist,
                       int32_t *hops, int32_t *visited, int32_t *list_ids, float *list_dist, int32_t *list_len,
                       int32_t *status, cudaStream_t s) {
    DR_CHECK(h->d_codes && h->d_codebook && h->M > 0, "dr_search: the u8-table mode needs PQ codes and the codebook");
    FastArgs a;
    memset(&a, 0, sizeof(a));
    a.vec = h->d_vec; a.adj = h->d_adj; a.codes = h->d_codes; a.deleted = p->ignore_deleted ? nullptr : h->d_deleted;
    a.N = h->N; a.D = h->D; a.R = h->R; a.M = h->M;
    a.W2 = p->w_after_empty > p->W ? p->w_after_empty : p->W;
    DR_CHECK(a.W2 <= 32, "dr_search: w_after_empty must be <= 32 (got %d)", p->w_after_empty);
    // (the next-in-line prefetch bits 2 / 8 / 16 assume a fixed W: cleared below when the doubling is on; they are off by default)
    a.k = p->k; a.L = p->L; a.W = p->W; a.rerank = p->rerank; a.sqrt_out = p->sqrt_out; a.prefetch = (a.W2 > p->W) ? (p->prefetch & ~(2 | 8 | 16)) : p->prefetch;
    if ((h->R & 3) != 0 || p->W > 16) a.prefetch &= ~16;   // bulk copies of adjacency rows need 16-byte rows; 16 barriers
    if (h->d_peer_recv) {
        DR_CHECK(p->rerank && !p->sqrt_out && B <= h->peer_B, "dr_search: the peer-routed exchange needs rerank = 1, sqrt_out = 0 and B <= the routed batch");
        a.peer_recv = h->d_peer_recv; a.peer_G = h->peer_G; a.peer_rank = h->peer_rank; a.peer_B = h->peer_B;
        a.peer_id_offset = h->peer_id_offset;
    }
    a.start = (uint32_t)(p->start_plus1 > 0 ? (int64_t)p->start_plus1 - 1 : h->medoid);   // range-checked by launch_search
    const bool word_layout = ((h->M & 3) == 0) && h->M <= 256;
    DR_CHECK(p->lut_fmt != DR_LUT_U8_TC || (word_layout && ((h->D / h->M) & 7) == 0),
             "dr_search: DR_LUT_U8_TC needs M %% 4 == 0, M <= 256 and (D / M) %% 8 == 0 (D=%d M=%d)", h->D, h->M);
    // hash_cap < 0: the visited set lives entirely in the CTA's global table (L2-resident, 32 KB per CTA): no shared-memory
    // hash, so a fourth CTA fits on the SM next to three 48 KB tables (the kernel is bound by resident queries)
    const bool l2_visited = p->hash_cap < 0 || (DR_L2V && p->hash_cap == 0);
    int shape = (h->R == 32 && p->W == 8 && a.W2 == DR_W2_SPEC) ? (h->D == 1536 ? 2 : 1) : 0;
    if (shape == 2 && p->L == 100 && p->hash_cap == 0) shape = 3;   // confirmed below once the table size is known
    fast_kernel_t kern = nullptr;
    // Region 0 holds the ADC table during the traversal and the rerank's row staging slots afterwards.  A small table
    // (M = 64: 16 KB) would leave room for only two 6 KB rows, i.e. two warps reranking: grow the region towards one slot per
    // warp as long as three CTAs still fit on the SM.
    const int slot_bytes = (h->D * 4 + 15) & ~15;
    int region0 = ((h->M * 256 + 15) / 16) * 16;
    if (p->rerank && (h->D & 3) == 0) {
        const int rest = 6 * 1024 + 4096 * 4 + 2048;          // lists, newcomers, a 4096-slot visited table, static + reserve
        for (int slots = DR_FAST_NT / 32; slots >= 1; --slots) {
            const int want = slots * slot_bytes;
            if (want <= region0) break;
            if (want + rest <= h->smem_optin / 3) { region0 = want; break; }
        }
    }
    int off = region0;
    const int LC = (p->L + 2) & ~1;
    a.o_list0 = off; off += LC * 8;
    a.o_list1 = off; off += LC * 8;
    a.o_ur0 = off; off += ((LC + 2) * 2 + 7) / 8 * 8;
    a.o_ur1 = off; off += ((LC + 2) * 2 + 7) / 8 * 8;
    const int NC = (a.W2 * h->R + 1) & ~1;
    a.o_newk = off; off += NC * 8;
    a.o_newid = off; off += NC * 4;
    a.o_sel = off; off += ((((p->W + a.W2) > DR_SELCAP ? (p->W + a.W2) : DR_SELCAP) * 4 + 7) / 8) * 8;
    off = (off + 15) / 16 * 16;
    a.o_adjrow = off;
    if (a.prefetch & 16) off += (p->W * h->R * 4 + 15) / 16 * 16;
    a.o_hash = off;
    const int fixed = off;
    // The query vector (rerank only) lives in the visited table's bytes once the traversal is over, behind the rerank
    // keys; a table too small for that (tests force small ones) gets a region of its own after it.
    const int q_in_hash = ((p->L * 8 + 15) / 16) * 16;
    const int q_bytes = p->rerank ? ((h->D * 4 + 15) / 16) * 16 : 0;
    // visited table: enough for the typical visit count at <= 3/4 load, then whatever keeps 3 CTAs per SM
    uint32_t hc = 0;
    int min_hash = 1024;
    while (min_hash * 4 < p->L * 8) min_hash <<= 1;
    a.rr_slots = region0 / slot_bytes;
    int q_extra = 0;
    if (l2_visited) {
        // rerank keys reuse the newcomer keys' bytes, the query vector takes the last staging slot of the table region
        a.o_rrk = a.o_newk;
        if (p->L > NC) { a.o_rrk = fixed; q_extra += ((p->L * 8 + 15) / 16) * 16; }
        if (q_bytes) {
            if (a.rr_slots >= 2) { a.rr_slots -= 1; a.o_q = a.rr_slots * slot_bytes; }
            else { a.o_q = fixed + q_extra; q_extra += q_bytes; a.rr_slots = 0; }   // rows come straight from global memory
        }
    } else if (p->hash_cap > 0) hc = (uint32_t)p->hash_cap;
    else {
        uint32_t want = 1024;
        long long target = (long long)(p->L + 2 * a.W2) * h->R;
        while ((long long)want * 3 / 4 < target && want < 65536) want <<= 1;
        hc = want;
        // prefer three CTAs per SM: shrink the table down to 4096 slots to get there (the overflow table takes the tail)
        while (hc > 4096 && fixed + (int)hc * 4 + 1024 > h->smem_optin / 3) hc >>= 1;
        while ((int)hc > min_hash && fixed + (int)hc * 4 + 256 > h->smem_optin) hc >>= 1;
    }
    if (!l2_visited) {
        a.o_rrk = a.o_hash;
        if (q_bytes && (int)hc * 4 < q_in_hash + q_bytes) { a.o_q = fixed + (int)hc * 4; q_extra = q_bytes; }
        else a.o_q = a.o_hash + q_in_hash;
        DR_CHECK((hc & (hc - 1)) == 0 && hc >= 64 && (int)hc * 4 >= p->L * 8 && fixed + (int)hc * 4 + q_extra + 512 <= h->smem_optin,
                 "dr_search(u8): visited table of %u slots is not usable (power of two, >= 64, >= 2L, %d B of shared memory needed)",
                 hc, fixed + (int)hc * 4 + q_extra);
    } else {
        DR_CHECK(fixed + q_extra + 512 <= h->smem_optin, "dr_search(u8): %d B of shared memory needed", fixed + q_extra);
    }
    a.hash_cap = hc;
    if (shape == 3 && hc != (DR_L2V ? 0u : 4096u)) shape = 2;
    if (shape == 3 && p->prefetch == DR_PF_SPEC && !a.deleted && p->rerank) shape = 4;
    kern = pick_fast_kernel(h->M, l2_visited ? 4 : 3, shape);
    const int smem = fixed + (int)hc * 4 + q_extra;
    const int nt = DR_FAST_NT;
    DR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    int occ = 0;
    DR_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, nt, smem));
    DR_CHECK(occ >= 1, "dr_search(u8): kernel does not fit (smem %d)", smem);
    const int max_grid = h->sms * occ;
    a.ovf_cap = 65536;
    if (l2_visited) {   // sized for the visit count (<= 1/2 load), small enough that all CTAs' tables stay in L2
        uint32_t want = 8192;
        while ((long long)want / 2 < (long long)(p->L + 2 * a.W2) * h->R && want < 65536) want <<= 1;
        a.ovf_cap = want;
    }
    size_t need = (size_t)max_grid * a.ovf_cap * 4;
    if (h->ovf_bytes < need) {
        if (h->d_ovf) cudaFree(h->d_ovf);
        h->d_ovf = nullptr; h->ovf_bytes = 0;
        DR_CUDA(cudaMalloc(&h->d_ovf, need));
        h->ovf_bytes = need;
        DR_CUDA(cudaMemsetAsync(h->d_ovf, 0xFF, need, s));
    }
    a.ovf = h->d_ovf;
    if (!h->d_counter) DR_CUDA(cudaMalloc(&h->d_counter, 16 * sizeof(u64)));
    a.counter = h->d_counter;

    const size_t per_q = (size_t)h->M * 256;
    int64_t cap = (int64_t)((size_t)8 << 30) / (int64_t)per_q;   // table scratch <= 8 GB per launch (160k queries at M = 192)
    if (p->chunk > 0) cap = p->chunk;
    if (cap < 1) cap = 1;
    const int64_t nchunks = (B + cap - 1) / cap;
    const int64_t chunk = (B + nchunks - 1) / nchunks;
    // scratch: table | scale | offset | mn | range
    const size_t sz_tab = (size_t)chunk * per_q;
    const size_t sz_f = ((size_t)chunk * 4 + 255) / 256 * 256;
    const size_t sz_mn = ((size_t)chunk * h->M * 4 + 255) / 256 * 256;
    if (dr_scratch((void **)&h->d_lut, &h->lut_bytes, sz_tab + 3 * sz_f + sz_mn)) return 1;
    uint8_t *d_tab = reinterpret_cast<uint8_t *>(h->d_lut);
    float *d_scale = reinterpret_cast<float *>(d_tab + sz_tab);
    float *d_off = reinterpret_cast<float *>(d_tab + sz_tab + sz_f);
    unsigned *d_range = reinterpret_cast<unsigned *>(d_tab + sz_tab + 2 * sz_f);
    float *d_mn = reinterpret_cast<float *>(d_tab + sz_tab + 3 * sz_f);

    for (int64_t c0 = 0; c0 < B; c0 += chunk) {
        const int64_t cb = (B - c0 < chunk) ? (B - c0) : chunk;
        if (p->lut_fmt == DR_LUT_U8_TC) {
            if (launch_lut_build_u8_tc(h->d_codebook, d_Q + (size_t)c0 * h->D, cb, h->D, h->M, d_tab, d_scale, d_off, d_mn, d_range, h->sms, s))
                return 1;
        } else if (launch_lut_build_u8(h->d_codebook, d_Q + (size_t)c0 * h->D, cb, h->D, h->M, d_tab, d_scale, d_off, d_mn, d_range,
                                       word_layout ? 1 : 0, s)) {
            return 1;
        }
        a.Q = d_Q + (size_t)c0 * h->D; a.lut8 = d_tab; a.lut_scale = d_scale; a.lut_offset = d_off; a.B = cb;
        a.out_ids = ids + (size_t)c0 * p->k;
        a.peer_q0 = c0;
        a.out_dist = dist ? dist + (size_t)c0 * p->k : nullptr;
        a.out_hops = hops ? hops + c0 : nullptr;
        a.out_visited = visited ? visited + c0 : nullptr;
        a.list_ids = list_ids ? list_ids + (size_t)c0 * p->L : nullptr;
        a.list_dist = list_dist ? list_dist + (size_t)c0 * p->L : nullptr;
        a.list_len = list_len ? list_len + c0 : nullptr;
        a.status = status ? status + c0 : nullptr;
#ifdef DR_PHASE_TIMING
        DR_CUDA(cudaMemsetAsync(h->d_counter, 0, 16 * sizeof(u64), s));
#else
        DR_CUDA(cudaMemsetAsync(h->d_counter, 0, sizeof(u64), s));
#endif
        int grid = (int)((cb < (int64_t)max_grid) ? cb : max_grid);
        if (h->timing) DR_CUDA(cudaEventRecord(h->ev0, s));
        kern<<<grid, nt, smem, s>>>(a);
        DR_LAUNCHED();
#ifdef DR_PHASE_TIMING
        {
            u64 hc[16];
            DR_CUDA(cudaStreamSynchronize(s));
            DR_CUDA(cudaMemcpy(hc, h->d_counter, sizeof(hc), cudaMemcpyDeviceToHost));
            static const char *nm[12] = {"fetch+output", "staging", "P1 visited+barrier", "P2 rest+barrier", "merge", "rerank", "P1 scan", "P1 adj wait", "P2 first group", "-", "-", "-"};
            double tot = 0;
            for (int i = 0; i < 9; ++i) tot += (double)hc[1 + i];
            fprintf(stderr, "[phase] %lld queries, grid %d:", (long long)cb, grid);
            for (int i = 0; i < 9; ++i) fprintf(stderr, " %s %.1f%% (%.0f cyc/query)", nm[i], 100.0 * hc[1 + i] / tot, (double)hc[1 + i] / (double)cb);
            fprintf(stderr, "\n");
        }
#endif
        if (h->timing) {
            DR_CUDA(cudaEventRecord(h->ev1, s));
            DR_CUDA(cudaEventSynchronize(h->ev1));
            float ms = 0.f;
            DR_CUDA(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
            h->timed_ms += ms;
            h->timed_launches += 1;
        }
    }
    return 0;
}
