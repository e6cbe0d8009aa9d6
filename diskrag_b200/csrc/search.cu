// search.cu — K1: batched best-first ("beam") search over the Vamana graph, one CTA per query.
//
// Replaces greedy_search_cython (cython_utils.pyx:72-122) + compute_query_distance
// (vamana_graph.py:301-329), greedy_search (vamana_graph.py:607-640), beam_search_from_disk
// (vamana_graph.py:719-760) and the traversal/exact-distance part of
// SearchEngineCorrect._pq_accelerated_graph_search (search_engine.py:398-506).
//
// Formulation (restated and pinned in oracle/oracle.c:orc_search_list): a sorted list of
// the L best (dist, id) keys seen; every step the first W unexpanded entries are expanded, their
// adjacency rows are scanned in stored order, first-seen ids enter the visited set and get a
// distance, the step's newcomers are merged and the best L stay.  W == 1 with strict ties is the
// reference's two-heap algorithm exactly, including its behaviour on exact distance ties (sequential
// strict-< accept, evict (max d, min id), and "ghost" expansion of an evicted node that ties with the
// worst kept distance).
//
// Per-CTA shared memory: the query's ADC table (M x 256 fp32), the query vector, two list buffers,
// the step's newcomer keys, the visited hash.  Gathers: adjacency rows with one coalesced warp load,
// PQ code rows with 128-bit loads, full vectors (exact mode / rerank) with coalesced float4 loads.
#include "common.cuh"

#ifndef DR_COMPACT_NEWK
#define DR_COMPACT_NEWK 1   // W > 1: survivors appended compactly (merge cost ~ survivors, not newcomers)
#endif
#ifndef DR_EXACT_X2
#define DR_EXACT_X2 1   // exact-distance phase of search_kernel: two rows per warp in flight (the graph build's search)
#endif

#include <vector>

struct SearchArgs {
    const float *vec; const uint32_t *adj; const uint8_t *codes;
    const int32_t *deg;   // optional true row lengths (graph under construction); NULL = rows are R long, 0-padded
    const uint8_t *deleted;  // optional lazy-delete mask: masked ids are skipped exactly like the reference's is_deleted test
    const int32_t *qmap;  // optional: query b is row qmap[b] of Q (build: queries are dataset points)
    const float *Q; const float *lut;
    long long N; int D, R, M;
    long long B; int k, L, W;
    int dist, adc_tree, rerank, sqrt_out, strict;
    uint32_t start;
    int32_t *out_ids; float *out_dist; int32_t *out_hops; int32_t *out_visited;
    int32_t *list_ids; float *list_dist; int32_t *list_len;
    int32_t *trace; int trace_cap; int32_t *status;
    u64 *counter;
    uint32_t *ovf; uint32_t ovf_cap;  // per-CTA global overflow table, ovf_cap slots each (power of two)
    uint32_t hash_cap;                // shared-memory visited table slots (power of two)
    int o_q, o_list0, o_list1, o_newk, o_newid, o_sel, o_ghost, o_hash;  // smem byte offsets (LUT at 0)
};

#define DR_MAX_GHOST 16

__device__ __forceinline__ float adc_seq(const uint8_t *__restrict__ code, const float *__restrict__ lut, int M) {
    float acc = 0.0f;
    if ((M & 15) == 0) {
        const uint4 *p = reinterpret_cast<const uint4 *>(code);
        const int nv = M >> 4;
#pragma unroll 2
        for (int c = 0; c < nv; ++c) {
            uint4 v = __ldg(p + c);
            const float *l = lut + c * 16 * 256;
            uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                acc = __fadd_rn(acc, l[(j * 4 + 0) * 256 + (w[j] & 0xFFu)]);
                acc = __fadd_rn(acc, l[(j * 4 + 1) * 256 + ((w[j] >> 8) & 0xFFu)]);
                acc = __fadd_rn(acc, l[(j * 4 + 2) * 256 + ((w[j] >> 16) & 0xFFu)]);
                acc = __fadd_rn(acc, l[(j * 4 + 3) * 256 + (w[j] >> 24)]);
            }
        }
    } else {
        for (int m = 0; m < M; ++m) acc = __fadd_rn(acc, lut[m * 256 + __ldg(code + m)]);
    }
    return acc;
}

// throughput order (oracle.c:adc_tree): lane l sums m = l, l+32, ... then xor-butterfly
__device__ __forceinline__ float adc_tree_warp(const uint8_t *__restrict__ code, const float *__restrict__ lut, int M, int lane) {
    float acc = 0.0f;
    for (int m = lane; m < M; m += 32) acc = __fadd_rn(acc, lut[m * 256 + __ldg(code + m)]);
    return warp_sum_butterfly(acc);
}

// Same order, G code rows per warp at once: each lane fetches one (two) 32-bit words of every row first — all loads
// in flight together — and the byte a lane needs (m = lane + 32 j) is pulled from its owner lane by shuffle.
// Needs M % 4 == 0 and M <= 256.
template <int G>
__device__ __forceinline__ void adc_tree_group(const uint8_t *__restrict__ codes, int M, const float *__restrict__ lut,
                                               const uint32_t (&ids)[G], int cnt, int lane, float (&out)[G]) {
    const int words = M >> 2;
    uint32_t w0[G], w1[G];
#pragma unroll
    for (int g = 0; g < G; ++g) {
        w0[g] = 0u; w1[g] = 0u;
        if (g < cnt) {
            const uint32_t *cw = reinterpret_cast<const uint32_t *>(codes + (size_t)ids[g] * M);
            if (lane < words) w0[g] = __ldg(cw + lane);
            if (lane + 32 < words) w1[g] = __ldg(cw + lane + 32);
        }
    }
    const int nj = (M + 31) >> 5;
    const int sh = (lane & 3) << 3;
#pragma unroll
    for (int g = 0; g < G; ++g) {
        if (g < cnt) {  // cnt is warp-uniform
            float acc = 0.0f;
            for (int j = 0; j < nj; ++j) {
                const int src = (lane >> 2) + ((j & 3) << 3);
                const uint32_t word = __shfl_sync(DR_FULL, (j < 4) ? w0[g] : w1[g], src);
                const int m = lane + (j << 5);
                if (m < M) acc = __fadd_rn(acc, lut[m * 256 + ((word >> sh) & 0xFFu)]);
            }
            out[g] = warp_sum_butterfly(acc);
        }
    }
}

// The reference's per-neighbour accept/evict rule, literally, for the (rare) step in which an exact
// distance tie sits on the list boundary.  Runs on one thread.  (oracle.c: orc_search_list, strict branch)
__device__ __noinline__ void seq_insert_strict(u64 *lst, int *pn, int L, const u64 *newk, int nn, u64 *ghost, int *png,
                                               int *pstatus) {
    int n = *pn, ng = *png;
    for (int t = 0; t < nn; ++t) {
        u64 key = newk[t];
        if (key == DR_KEY_MAX) continue;
        if (n >= L && !(key_dbits(key) < key_dbits(lst[n - 1]))) continue;
        int pos = n;
        while (pos > 0 && key < lst[pos - 1]) { lst[pos] = lst[pos - 1]; --pos; }
        lst[pos] = key;
        ++n;
        if (n > L) {
            int e = n - 1;
            uint32_t wd = key_dbits(lst[n - 1]);
            while (e > 0 && key_dbits(lst[e - 1]) == wd) --e;
            u64 ev = lst[e];
            for (int i = e; i < n - 1; ++i) lst[i] = lst[i + 1];
            --n;
            if (!(ev & 1ull) && key_dbits(ev) == key_dbits(lst[n - 1])) {
                if (ng < DR_MAX_GHOST) ghost[ng++] = ev;
                else *pstatus |= DR_ST_TIE_OVERFLOW;
            }
        }
    }
    *pn = n;
    *png = ng;
}

extern __shared__ __align__(16) unsigned char dr_smem[];

__global__ void __launch_bounds__(512, 1) search_kernel(const SearchArgs a) {
    float *s_lut = reinterpret_cast<float *>(dr_smem);
    float *s_q = reinterpret_cast<float *>(dr_smem + a.o_q);
    u64 *s_list0 = reinterpret_cast<u64 *>(dr_smem + a.o_list0);
    u64 *s_list1 = reinterpret_cast<u64 *>(dr_smem + a.o_list1);
    u64 *s_newk = reinterpret_cast<u64 *>(dr_smem + a.o_newk);
    uint32_t *s_newid = reinterpret_cast<uint32_t *>(dr_smem + a.o_newid);
    uint32_t *s_sel = reinterpret_cast<uint32_t *>(dr_smem + a.o_sel);      // [W] ids, then [W] list indices
    u64 *s_ghost = reinterpret_cast<u64 *>(dr_smem + a.o_ghost);
    uint32_t *s_hash = reinterpret_cast<uint32_t *>(dr_smem + a.o_hash);
    u64 *s_rrk = reinterpret_cast<u64 *>(dr_smem + a.o_hash);  // rerank keys alias the (dead) visited table

    __shared__ long long s_b;
    __shared__ __align__(8) uint64_t s_lutbar;
    __shared__ int s_n, s_nn, s_ns, s_ng, s_mvalid, s_hcount, s_ovfcount, s_useovf, s_ovfused, s_status;

    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, wid = tid >> 5, nw = nt >> 5;
    const uint32_t lt_mask = (1u << lane) - 1u;
    const int D = a.D, R = a.R, M = a.M, L = a.L, W = a.W;
    const uint32_t hmask = a.hash_cap - 1u, ovf_mask = a.ovf_cap - 1u;
    const int hlimit = (int)(a.hash_cap - (a.hash_cap >> 2));       // 3/4 load
    const int ovf_limit = (int)(a.ovf_cap - (a.ovf_cap >> 2));
    uint32_t *my_ovf = a.ovf + (size_t)blockIdx.x * a.ovf_cap;
    const bool pq = (a.dist == DR_DIST_PQ);
    const bool cosine = (a.dist == DR_DIST_COSINE);
    const bool strict = a.strict != 0;
    uint32_t lut_phase = 0;
    if (tid == 0) { mbar_init(&s_lutbar, 1); fence_mbar_init(); }

    for (;;) {
        __syncthreads();
        if (tid == 0) s_b = (long long)atomicAdd(a.counter, 1ull);
        __syncthreads();
        const long long b = s_b;
        if (b >= a.B) break;

        // ---- stage the query, its ADC table, and clear the visited table --------------------------
        {
            const float *qg = a.Q + (size_t)(a.qmap ? a.qmap[b] : b) * D;
            for (int i = tid; i < D; i += nt) s_q[i] = __ldg(qg + i);
            if (pq && tid == 0) {  // one bulk async copy (TMA engine) of the query's M x 256 table
                mbar_expect_tx(&s_lutbar, (uint32_t)M * 1024u);
                bulk_g2s(s_lut, a.lut + (size_t)b * M * 256, (uint32_t)M * 1024u, &s_lutbar);
            }
            for (uint32_t i = tid; i < a.hash_cap; i += nt) s_hash[i] = DR_EMPTY;
            if (tid == 0) { s_ng = 0; s_hcount = 1; s_ovfcount = 0; s_ovfused = 0; s_status = 0; }
            if (pq) { mbar_wait(&s_lutbar, lut_phase); lut_phase ^= 1u; }
        }
        __syncthreads();

        // ---- start node ----------------------------------------------------------------------------
        if (wid == 0) {
            float d0;
            if (pq) {
                if (a.adc_tree) d0 = adc_tree_warp(a.codes + (size_t)a.start * M, s_lut, M, lane);
                else d0 = adc_seq(a.codes + (size_t)a.start * M, s_lut, M);  // every lane computes the same value
            } else {
                d0 = cosine ? warp_cosdist(a.vec + (size_t)a.start * D, s_q, D, lane) : warp_l2sq(a.vec + (size_t)a.start * D, s_q, D, lane);
            }
            if (lane == 0) {
                s_list0[0] = make_key(d0, a.start);
                s_hash[hash_u32(a.start) & hmask] = a.start;
                if (a.trace && a.trace_cap > 0) a.trace[(size_t)b * a.trace_cap] = (int32_t)a.start;
            }
        }
        int cur = 0, n = 1, hops = 0, nvis = 1;
        __syncthreads();

        // ---- traversal -----------------------------------------------------------------------------
        for (;;) {
            u64 *lst = cur ? s_list1 : s_list0;
            u64 *oth = cur ? s_list0 : s_list1;

            // (1) pick the first W unexpanded entries (key order)
            if (wid == 0) {
                int found = 0;
                for (int base = 0; base < n && found < W; base += 32) {
                    int i = base + lane;
                    u64 kx = (i < n) ? lst[i] : 1ull;
                    bool un = !(kx & 1ull);
                    unsigned m = __ballot_sync(DR_FULL, un);
                    int r = found + __popc(m & lt_mask);
                    if (un && r < W) { s_sel[r] = key_id(kx); s_sel[W + r] = (uint32_t)i; }
                    found += __popc(m);
                }
                if (found > W) found = W;
                __syncwarp();
                bool ghost_taken = false;
                const int ng_now = s_ng;   // every lane reads it before lane 0 may rewrite it
                __syncwarp();
                if (strict && ng_now > 0) {  // W == 1
                    if (lane == 0) {
                        int ng = ng_now;
                        if (key_dbits(s_ghost[0]) > key_dbits(lst[n - 1])) ng = 0;  // worst improved: ghosts are dead
                        if (ng > 0) {
                            int gb = 0;
                            for (int g = 1; g < ng; ++g) if (s_ghost[g] < s_ghost[gb]) gb = g;
                            if (found == 0 || s_ghost[gb] < lst[s_sel[W]]) {
                                s_sel[0] = key_id(s_ghost[gb]);
                                s_ghost[gb] = s_ghost[--ng];
                                ghost_taken = true;
                            }
                        }
                        s_ng = ng;
                    }
                    ghost_taken = __shfl_sync(DR_FULL, (int)ghost_taken, 0) != 0;
                    if (ghost_taken) found = 1;
                }
                if (!ghost_taken && lane < found) lst[s_sel[W + lane]] |= 1ull;  // W <= 32
                if (lane == 0) {
                    s_ns = found;
                    s_nn = 0;
                    s_mvalid = 0;
                    int useovf = (s_hcount + W * R > hlimit) ? 1 : 0;
                    s_useovf = useovf;
                    if (useovf) {
                        s_ovfused = 1;
                        if (s_ovfcount + W * R > ovf_limit) { s_status |= DR_ST_VISITED_OVERFLOW; s_ns = 0; }
                    }
                }
            }
            __syncthreads();
            const int ns = s_ns;
            if (ns == 0) break;
            const bool use_ovf = s_useovf != 0;

            // (2) adjacency rows -> first-seen neighbours, in stored order (one warp per expanded node)
            for (int s = wid; s < ns; s += nw) {
                const uint32_t *row = a.adj + (size_t)s_sel[s] * R;
                const int len = a.deg ? a.deg[s_sel[s]] : R;
                for (int j0 = 0; j0 < len; j0 += 32) {
                    int j = j0 + lane;
                    uint32_t nb = (j < len) ? row[j] : DR_EMPTY;
                    bool valid = (j < len) && ((long long)nb < a.N);
                    if (valid && a.deleted) valid = a.deleted[nb] == 0;
                    unsigned peers = __match_any_sync(DR_FULL, nb);
                    bool leader = (__ffs(peers) - 1) == lane;  // 0-padding repeats an id inside a row
                    bool isnew = false;
                    if (valid && leader) isnew = visited_insert(nb, s_hash, hmask, use_ovf, my_ovf, ovf_mask);
                    unsigned m = __ballot_sync(DR_FULL, isnew);
                    int cnt = __popc(m);
                    int basepos = 0;
                    if (cnt) {
                        if (lane == 0) basepos = atomicAdd(&s_nn, cnt);
                        basepos = __shfl_sync(DR_FULL, basepos, 0);
                    }
                    if (isnew) s_newid[basepos + __popc(m & lt_mask)] = nb;
                }
            }
            __syncthreads();

            // (3) distances of the newcomers; drop those that cannot enter a full list
            const int nn = s_nn;
            const bool full = (n >= L);
            const u64 worstk = lst[n - 1] & ~1ull;
            const uint32_t worst_db = key_dbits(worstk);
            // survivors are appended compactly when their order cannot matter (keys are unique and the rank merge is
            // order-free): the merge then costs (n + survivors) x survivors compares instead of (n + newcomers) x newcomers.
            // The reference-order tie mode keeps one slot per newcomer: its sequential fallback inserts in discovery order.
            const bool compact = DR_COMPACT_NEWK && !strict;
            if (pq && !a.adc_tree) {
                for (int i = tid; i < nn; i += nt) {
                    uint32_t id = s_newid[i];
                    float d = adc_seq(a.codes + (size_t)id * M, s_lut, M);
                    u64 key = make_key(d, id);
                    bool ok = !full || (strict ? (key_dbits(key) < worst_db) : (key < worstk));
                    if (compact) { if (ok) s_newk[atomicAdd(&s_mvalid, 1)] = key; }
                    else { s_newk[i] = ok ? key : DR_KEY_MAX; if (ok) atomicAdd(&s_mvalid, 1); }
                }
            } else if (pq && (M & 3) == 0 && M <= 256) {
                constexpr int G = 4;
                for (int base = wid; base < nn; base += nw * G) {
                    uint32_t gid[G];
                    float gd[G];
                    int cnt = 0;
#pragma unroll
                    for (int g = 0; g < G; ++g) {
                        const int i = base + g * nw;
                        gid[g] = 0u;
                        if (i < nn) { gid[g] = s_newid[i]; cnt = g + 1; }
                    }
                    adc_tree_group<G>(a.codes, M, s_lut, gid, cnt, lane, gd);
                    if (lane == 0) {
#pragma unroll
                        for (int g = 0; g < G; ++g) {
                            if (g < cnt) {
                                u64 key = make_key(gd[g], gid[g]);
                                bool ok = !full || (strict ? (key_dbits(key) < worst_db) : (key < worstk));
                                if (compact) { if (ok) s_newk[atomicAdd(&s_mvalid, 1)] = key; }
                                else { s_newk[base + g * nw] = ok ? key : DR_KEY_MAX; if (ok) atomicAdd(&s_mvalid, 1); }
                            }
                        }
                    }
                }
            } else if (!pq && !cosine && DR_EXACT_X2) {
                // exact L2: two rows per warp at a time (twice the bytes in flight per warp; each row keeps the canonical order)
                for (int i = wid; i < nn; i += 2 * nw) {
                    const int i2 = i + nw;
                    const uint32_t idA = s_newid[i], idB = s_newid[i2 < nn ? i2 : i];
                    float dA, dB;
                    if (i2 < nn) warp_l2sq_x2(a.vec + (size_t)idA * D, a.vec + (size_t)idB * D, s_q, D, lane, dA, dB);
                    else { dA = warp_l2sq(a.vec + (size_t)idA * D, s_q, D, lane); dB = 0.0f; }
                    if (lane == 0) {
                        const u64 keyA = make_key(dA, idA);
                        const bool okA = !full || (strict ? (key_dbits(keyA) < worst_db) : (keyA < worstk));
                        const u64 keyB = make_key(dB, idB);
                        const bool okB = i2 < nn && (!full || (strict ? (key_dbits(keyB) < worst_db) : (keyB < worstk)));
                        const int okn = (okA ? 1 : 0) + (okB ? 1 : 0);
                        int slot = okn ? atomicAdd(&s_mvalid, okn) : 0;
                        if (compact) {
                            if (okA) s_newk[slot++] = keyA;
                            if (okB) s_newk[slot] = keyB;
                        } else {
                            s_newk[i] = okA ? keyA : DR_KEY_MAX;
                            if (i2 < nn) s_newk[i2] = okB ? keyB : DR_KEY_MAX;
                        }
                    }
                }
            } else {
                for (int i = wid; i < nn; i += nw) {
                    uint32_t id = s_newid[i];
                    float d = pq ? adc_tree_warp(a.codes + (size_t)id * M, s_lut, M, lane)
                                 : (cosine ? warp_cosdist(a.vec + (size_t)id * D, s_q, D, lane) : warp_l2sq(a.vec + (size_t)id * D, s_q, D, lane));
                    if (lane == 0) {
                        u64 key = make_key(d, id);
                        bool ok = !full || (strict ? (key_dbits(key) < worst_db) : (key < worstk));
                        if (compact) { if (ok) s_newk[atomicAdd(&s_mvalid, 1)] = key; }
                        else { s_newk[i] = ok ? key : DR_KEY_MAX; if (ok) atomicAdd(&s_mvalid, 1); }
                    }
                }
            }
            if (a.trace) {
                for (int i = tid; i < nn; i += nt)
                    if (nvis + i < a.trace_cap) a.trace[(size_t)b * a.trace_cap + nvis + i] = (int32_t)s_newid[i];
            }
            if (tid == 0) {
                if (use_ovf) s_ovfcount += nn; else s_hcount += nn;
            }
            __syncthreads();
            nvis += nn;
            hops += ns;
            const int mvalid = s_mvalid;

            // (4) rank-merge newcomers into the other buffer; best L stay
            if (mvalid > 0) {
                const int nk = compact ? mvalid : nn;      // keys in s_newk: the survivors only, or one slot per newcomer
                const int total = n + nk;
                for (int x = tid; x < total; x += nt) {
                    u64 key;
                    int pos;
                    if (x < n) {
                        key = lst[x];
                        int c = 0;
                        for (int j = 0; j < nk; ++j) c += (s_newk[j] < key) ? 1 : 0;
                        pos = x + c;
                    } else {
                        key = s_newk[x - n];
                        if (key == DR_KEY_MAX) continue;
                        int c = 0;
                        for (int j = 0; j < nk; ++j) c += (s_newk[j] < key) ? 1 : 0;
                        int lo = 0, hi = n;
                        while (lo < hi) {
                            int mid = (lo + hi) >> 1;
                            if (lst[mid] < key) lo = mid + 1; else hi = mid;
                        }
                        pos = lo + c;
                    }
                    if (pos <= L) oth[pos] = key;  // slot L keeps the first dropped key (tie detection)
                }
                __syncthreads();
                const int tot = n + mvalid;
                const bool hazard = strict && tot > L && key_dbits(oth[L - 1]) == key_dbits(oth[L]);
                if (!hazard) {
                    cur ^= 1;
                    n = tot < L ? tot : L;
                } else {
                    if (tid == 0) {
                        int nn_ = n, ng_ = s_ng, st_ = s_status;
                        seq_insert_strict(lst, &nn_, L, s_newk, nn, s_ghost, &ng_, &st_);
                        s_n = nn_; s_ng = ng_; s_status = st_;
                    }
                    __syncthreads();
                    n = s_n;
                }
            } else {
                __syncthreads();   // nothing survived: still separate this step's reads of the step counters from the next selection
            }
        }

        // ---- results ---------------------------------------------------------------------------------
        const u64 *lst = cur ? s_list1 : s_list0;
        if (a.list_ids) {
            for (int i = tid; i < L; i += nt) {
                a.list_ids[(size_t)b * L + i] = i < n ? (int32_t)key_id(lst[i]) : -1;
                if (a.list_dist) a.list_dist[(size_t)b * L + i] = i < n ? key_dist(lst[i]) : __int_as_float(0x7f800000);
            }
        }
        const int k = a.k;
        if (a.rerank) {
            // exact fp32 L2^2 of every list entry (search_engine.py:374-379), stable sort, first k
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
            __syncthreads();
            for (int i = tid; i < n; i += nt) {
                u64 key = s_rrk[i];
                int pos = 0;
                for (int j = 0; j < n; ++j) pos += (s_rrk[j] < key) ? 1 : 0;
                if (pos < k) {
                    float d2 = ord2f((uint32_t)(key >> 32));
                    a.out_ids[(size_t)b * k + pos] = (int32_t)key_id(lst[i]);
                    if (a.out_dist) a.out_dist[(size_t)b * k + pos] = a.sqrt_out ? sqrtf(d2) : d2;
                }
            }
        } else {
            for (int i = tid; i < n && i < k; i += nt) {
                float d = key_dist(lst[i]);
                a.out_ids[(size_t)b * k + i] = (int32_t)key_id(lst[i]);
                if (a.out_dist) a.out_dist[(size_t)b * k + i] = a.sqrt_out ? sqrtf(d) : d;
            }
        }
        for (int i = n + tid; i < k; i += nt) {
            a.out_ids[(size_t)b * k + i] = -1;
            if (a.out_dist) a.out_dist[(size_t)b * k + i] = __int_as_float(0x7f800000);
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
}


// ---------------------------------------------------------------------------------------------------
// host launcher
// ---------------------------------------------------------------------------------------------------
static inline int align_up(int x, int a) { return (x + a - 1) / a * a; }

static int launch_search_impl(dr_index *h, const float *d_Q, int64_t B, const dr_search_params *p, const float *d_lut,
                              int32_t *ids, float *dist, int32_t *hops, int32_t *visited, int32_t *list_ids, float *list_dist,
                              int32_t *list_len, int32_t *trace, int32_t trace_cap, int32_t *status, cudaStream_t s,
                              const int32_t *qmap);

// The handle's scratch (ADC tables, work counter, overflow tables) is shared by every call on the handle.  Calls enqueued on
// ONE stream are ordered by the stream; a call on a different stream first waits, on the device, for the previous call's
// kernels (an event recorded after every launch), so two in-flight calls can never share the scratch.
int launch_search(dr_index *h, const float *d_Q, int64_t B, const dr_search_params *p, const float *d_lut,
                  int32_t *ids, float *dist, int32_t *hops, int32_t *visited, int32_t *list_ids, float *list_dist,
                  int32_t *list_len, int32_t *trace, int32_t trace_cap, int32_t *status, cudaStream_t s,
                  const int32_t *qmap) {
    if (h->ev_scratch && h->scratch_used && h->scratch_stream != s) DR_CUDA(cudaStreamWaitEvent(s, h->ev_scratch, 0));
    const int rc = launch_search_impl(h, d_Q, B, p, d_lut, ids, dist, hops, visited, list_ids, list_dist, list_len, trace, trace_cap,
                                      status, s, qmap);
    if (h->ev_scratch && B > 0) {
        DR_CUDA(cudaEventRecord(h->ev_scratch, s));
        h->scratch_stream = s; h->scratch_used = true;
    }
    return rc;
}

static int launch_search_impl(dr_index *h, const float *d_Q, int64_t B, const dr_search_params *p, const float *d_lut,
                              int32_t *ids, float *dist, int32_t *hops, int32_t *visited, int32_t *list_ids, float *list_dist,
                              int32_t *list_len, int32_t *trace, int32_t trace_cap, int32_t *status, cudaStream_t s,
                              const int32_t *qmap) {
    DR_CHECK(p->k >= 1 && p->L >= 1 && p->L <= 512 && p->k <= p->L, "dr_search: need 1 <= k <= L <= 512 (k=%d L=%d)", p->k, p->L);
    DR_CHECK(p->W >= 1 && p->W <= 32, "dr_search: W must be in 1..32 (got %d)", p->W);
    const bool pq = p->dist == DR_DIST_PQ;
    DR_CHECK(p->dist == DR_DIST_PQ || p->dist == DR_DIST_EXACT || p->dist == DR_DIST_COSINE, "dr_search: unknown dist %d", p->dist);
    DR_CHECK(p->dist != DR_DIST_COSINE || !p->rerank, "dr_search: the rerank is a squared-L2 rerank; DR_DIST_COSINE returns the traversal's own order");
    if (pq) DR_CHECK(h->d_codes && h->M > 0, "dr_search: index has no PQ codes");
    const int64_t start = p->start_plus1 > 0 ? (int64_t)p->start_plus1 - 1 : h->medoid;
    DR_CHECK(p->start_plus1 >= 0 && start >= 0 && start < h->N, "dr_search: start node %lld out of range (N=%lld)", (long long)start,
             (long long)h->N);
    if (B == 0) return 0;
    if (p->lut_fmt == DR_LUT_U8 || p->lut_fmt == DR_LUT_U8_TC) {
        DR_CHECK(pq && !d_lut && !trace && !qmap && !h->d_deg, "dr_search: the u8-table mode is PQ-only, builds its own tables and has no trace");
        return launch_search_fast(h, d_Q, B, p, ids, dist, hops, visited, list_ids, list_dist, list_len, status, s);
    }

    SearchArgs a;
    memset(&a, 0, sizeof(a));
    a.vec = h->d_vec; a.adj = h->d_adj; a.codes = h->d_codes;
    a.deg = h->d_deg;
    a.deleted = p->ignore_deleted ? nullptr : h->d_deleted;
    a.N = h->N; a.D = h->D; a.R = h->R; a.M = h->M;
    a.k = p->k; a.L = p->L; a.W = p->W;
    a.dist = p->dist; a.adc_tree = (pq && p->adc_order == DR_ADC_TREE) ? 1 : 0;
    a.rerank = p->rerank; a.sqrt_out = p->sqrt_out;
    a.strict = (p->W == 1) ? 1 : 0;
    a.start = (uint32_t)start;
    a.trace_cap = trace ? trace_cap : 0;

    // shared-memory layout
    int off = pq ? h->M * 1024 : 0;
    a.o_q = off; off += align_up(h->D * 4, 16);
    const int LC = align_up(p->L + 1, 2);
    a.o_list0 = off; off += LC * 8;
    a.o_list1 = off; off += LC * 8;
    const int NC = align_up(p->W * h->R, 2);
    a.o_newk = off; off += NC * 8;
    a.o_newid = off; off += NC * 4;
    a.o_sel = off; off += align_up(2 * p->W * 4, 8);
    a.o_ghost = off; off += DR_MAX_GHOST * 8;
    a.o_hash = off;
    const int fixed = off;
    int avail = h->smem_optin - fixed - 256;  // static __shared__ scalars
    int min_hash = 1024;
    while (min_hash * 4 < p->L * 8) min_hash <<= 1;  // rerank keys alias the table
    DR_CHECK(avail >= min_hash * 4, "dr_search: M=%d D=%d L=%d W=%d R=%d need %d B of shared memory, device offers %d",
             h->M, h->D, p->L, p->W, h->R, fixed + min_hash * 4 + 256, h->smem_optin);
    uint32_t hc;
    if (p->hash_cap > 0) {
        hc = (uint32_t)p->hash_cap;
        DR_CHECK((hc & (hc - 1)) == 0 && (int)hc * 4 <= avail && (int)hc * 4 >= p->L * 8 && hc >= 64,
                 "dr_search: hash_cap must be a power of two, >= 64, >= 2L and fit shared memory");
    } else {
        // enough for the typical visited count (~ (L+2) * R * 0.6) at <= 3/4 load, capped by what is left
        uint32_t want = 1024;
        long long target = (long long)(p->L + 2 * p->W) * h->R;
        while ((long long)want * 3 / 4 < target && want < 65536) want <<= 1;
        hc = want;
        while ((int)hc * 4 > avail) hc >>= 1;
    }
    a.hash_cap = hc;
    const int smem = fixed + (int)hc * 4;

    // one CTA per SM when the table is big: give it 16 warps; small tables keep 8 warps and co-reside
    int nt = p->threads > 0 ? p->threads : (smem > h->smem_optin / 2 ? 512 : 256);
    DR_CHECK(nt % 32 == 0 && nt >= 32 && nt <= 512, "dr_search: threads must be a multiple of 32 in 32..512");

    DR_CUDA(cudaFuncSetAttribute(search_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    int occ = 0;
    DR_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, search_kernel, nt, smem));
    DR_CHECK(occ >= 1, "dr_search: kernel does not fit (smem %d)", smem);
    const int max_grid = h->sms * occ;

    // per-CTA overflow table for the visited set
    a.ovf_cap = 65536;
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

    // chunking bounds the internal LUT buffer
    int64_t chunk = B;
    if (pq && !d_lut) {
        const size_t per_q = (size_t)h->M * 1024;
        int64_t cap = (int64_t)((size_t)4 << 30) / (int64_t)per_q;
        if (cap < 1) cap = 1;
        if (p->chunk > 0) cap = p->chunk;
        if (chunk > cap) chunk = cap;
        DR_CHECK(h->d_codebook, "dr_search: index has no codebook and no LUT was supplied");
        if (dr_scratch((void **)&h->d_lut, &h->lut_bytes, (size_t)chunk * per_q)) return 1;
    } else if (p->chunk > 0 && chunk > p->chunk) {
        chunk = p->chunk;
    }

    for (int64_t c0 = 0; c0 < B; c0 += chunk) {
        const int64_t cb = (B - c0 < chunk) ? (B - c0) : chunk;
        const float *lut_c = nullptr;
        if (pq) {
            if (d_lut) lut_c = d_lut + (size_t)c0 * h->M * 256;
            else {
                DR_CHECK(!qmap, "dr_search: qmap with an internal LUT is not supported");
                if (launch_lut_build(h->d_codebook, d_Q + (size_t)c0 * h->D, cb, h->D, h->M, h->d_lut, s)) return 1;
                lut_c = h->d_lut;
            }
        }
        a.Q = qmap ? d_Q : d_Q + (size_t)c0 * h->D; a.qmap = qmap ? qmap + c0 : nullptr; a.lut = lut_c; a.B = cb;
        a.out_ids = ids + (size_t)c0 * p->k;
        a.out_dist = dist ? dist + (size_t)c0 * p->k : nullptr;
        a.out_hops = hops ? hops + c0 : nullptr;
        a.out_visited = visited ? visited + c0 : nullptr;
        a.list_ids = list_ids ? list_ids + (size_t)c0 * p->L : nullptr;
        a.list_dist = list_dist ? list_dist + (size_t)c0 * p->L : nullptr;
        a.list_len = list_len ? list_len + c0 : nullptr;
        a.trace = trace ? trace + (size_t)c0 * trace_cap : nullptr;
        a.status = status ? status + c0 : nullptr;
        DR_CUDA(cudaMemsetAsync(h->d_counter, 0, sizeof(u64), s));
        int grid = (int)((cb < (int64_t)max_grid) ? cb : max_grid);
        if (h->timing) DR_CUDA(cudaEventRecord(h->ev0, s));
        search_kernel<<<grid, nt, smem, s>>>(a);
        DR_LAUNCHED();
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
