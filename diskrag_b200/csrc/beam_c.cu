// beam_c.cu — variant C with the reference's OWN semantics: beam_search_with_pq (vamana_graph.py:535-605) and beam_search
// (vamana_graph.py:690-717), the k-capped beam whose frontier truncation `while len(beam) > beam_width: heappop(beam)` (:595-596)
// drops the BEST frontier entries.  search.cu's L-list search is what the shims run by default (it returns better neighbours);
// this kernel exists so that a caller who depends on the reference's exact answers of that function can have them
// (shim: beam_search_with_pq(..., reference_semantics=True)).  Restated and pinned against the real reference in
// oracle/oracle.c:orc_beam_c (tests/test_oracle_vs_reference.py::test_variant_C_beam_search_equal).
//
// One warp per query (the reference's API is one query per call; this is a compatibility path, not the throughput path):
//   frontier  : u64 keys (dist, id) kept sorted DESCENDING in shared memory, so "pop the best" is `--n` and the reference's
//               truncation (keep the beam_width worst) is `n = min(n, beam_width)`;
//   top_k     : <= k keys, unsorted; eviction rule of a heapq of (-dist, id): the largest distance, smallest id among ties;
//   visited   : a bitmap over N in global memory, one per resident CTA, cleared per query;
//   a row     : 32 neighbours per pass; duplicates inside a pass resolved to the first occurrence (match_any), first-seen bits
//               claimed with atomicOr; distances for the newcomers (ADC: one lane per neighbour, sequential fp32 sum in table
//               order == asymmetric_distance_sq, fast_pq.py:320-328; exact: one warp per row in the canonical warp order);
//   accept    : the reference's per-neighbour rule (:589-593) applied by lane 0 in stored order.
// Every loop is bounded (a node is pushed at most once, so at most N pops).
#include "common.cuh"

struct BeamCArgs {
    const float *vec; const uint32_t *adj; const uint8_t *codes; const int32_t *deg; const uint8_t *deleted;
    const float *Q; const float *lut;
    long long N; int D, R, M;
    long long B; int k, bw, dist, sqrt_out;
    uint32_t start;
    int32_t *out_ids; float *out_dist; int32_t *out_hops; int32_t *out_visited;
    uint32_t *bitmap; long long words;   // per-CTA visited bitmap, `words` 32-bit words each
};

__device__ __forceinline__ float beamc_adc_seq(const uint8_t *__restrict__ code, const float *__restrict__ lut, int M) {
    float acc = 0.0f;
    for (int m = 0; m < M; ++m) acc = __fadd_rn(acc, __ldg(lut + m * 256 + __ldg(code + m)));
    return acc;
}

extern __shared__ __align__(16) unsigned char beamc_smem[];

__global__ void __launch_bounds__(32) beam_c_kernel(const BeamCArgs a) {
    const int lane = threadIdx.x;
    const int cap = a.bw + a.R + 1;
    float *q_s = reinterpret_cast<float *>(beamc_smem);                                   // [D] (exact mode only)
    u64 *beam = reinterpret_cast<u64 *>(beamc_smem + (size_t)((a.dist == DR_DIST_PQ ? 0 : a.D) * 4 + 15) / 16 * 16);  // [cap] descending
    u64 *top = beam + cap;                                                                // [k + 1]
    u64 *cand = top + a.k + 1;                                                            // [32] newcomers of a pass, stored order
    int *state = reinterpret_cast<int *>(cand + 32);                                      // nb, nt, worst dbits
    uint32_t *bm = a.bitmap + (size_t)blockIdx.x * a.words;

    for (long long b = blockIdx.x; b < a.B; b += gridDim.x) {
        for (long long w = lane; w < a.words; w += 32) bm[w] = 0u;
        const float *lut = a.lut ? a.lut + (size_t)b * a.M * 256 : nullptr;
        if (a.dist != DR_DIST_PQ)
            for (int i = lane; i < a.D; i += 32) q_s[i] = a.Q[(size_t)b * a.D + i];
        __syncwarp();

        // start node (:556, :569-570)
        float d0;
        if (a.dist == DR_DIST_PQ) d0 = beamc_adc_seq(a.codes + (size_t)a.start * a.M, lut, a.M);
        else d0 = warp_l2sq(a.vec + (size_t)a.start * a.D, q_s, a.D, lane);
        if (lane == 0) {
            bm[a.start >> 5] |= 1u << (a.start & 31);
            const u64 k0 = make_key(d0, a.start);
            beam[0] = k0; top[0] = k0;
            state[0] = 1; state[1] = 1; state[2] = (int)key_dbits(k0);
        }
        __syncwarp();
        int hops = 0, nvis = 1;

        for (long long it = 0; it <= a.N; ++it) {
            int nb = state[0];
            const int nt = state[1];
            const uint32_t worst = (uint32_t)state[2];
            if (nb == 0) break;
            const u64 cur = beam[nb - 1];                                  // heappop(beam): the smallest (dist, id)
            --nb;
            __syncwarp();
            if (lane == 0) state[0] = nb;
            __syncwarp();
            const uint32_t cid = key_id(cur);
            if (a.deleted && a.deleted[cid]) continue;                     // :577-578
            if (key_dbits(cur) > worst && nt == a.k) break;                // :580-581
            ++hops;
            const int rowlen = a.deg ? min(a.deg[cid], a.R) : a.R;
            const uint32_t *row = a.adj + (size_t)cid * a.R;
            for (int base = 0; base < rowlen; base += 32) {
                const int j = base + lane;
                uint32_t id = 0xFFFFFFFFu - (uint32_t)lane;                 // unique filler for lanes without a neighbour
                bool valid = false;
                if (j < rowlen) {
                    const uint32_t v = __ldg(row + j);
                    if ((long long)v < a.N && !(a.deleted && a.deleted[v])) { id = v; valid = true; }   // :584
                }
                const uint32_t same = __match_any_sync(DR_FULL, id);
                bool fresh = false;
                if (valid && (__ffs(same) - 1) == lane) {                  // first occurrence inside this pass
                    const uint32_t bit = 1u << (id & 31);
                    fresh = (atomicOr(&bm[id >> 5], bit) & bit) == 0u;
                }
                const uint32_t newmask = __ballot_sync(DR_FULL, fresh);
                if (newmask == 0u) continue;
                float d = 0.0f;
                if (a.dist == DR_DIST_PQ) {
                    if (fresh) d = beamc_adc_seq(a.codes + (size_t)id * a.M, lut, a.M);
                } else {
                    for (uint32_t mm = newmask; mm; mm &= mm - 1) {
                        const int src = __ffs(mm) - 1;
                        const uint32_t sid = __shfl_sync(DR_FULL, id, src);
                        const float ds = warp_l2sq(a.vec + (size_t)sid * a.D, q_s, a.D, lane);
                        if (lane == src) d = ds;
                    }
                }
                if (fresh) cand[__popc(newmask & ((1u << lane) - 1u))] = make_key(d, id);
                __syncwarp();
                const int nn = __popc(newmask);
                nvis += nn;
                if (lane == 0) {
                    int n_b = state[0], n_t = state[1];
                    uint32_t wd = (uint32_t)state[2];
                    for (int t = 0; t < nn; ++t) {
                        const u64 key = cand[t];
                        if (!(n_t < a.k || key_dbits(key) < wd)) continue;                // :589
                        int pos = n_b;                                                   // heappush(beam): keep it descending
                        while (pos > 0 && beam[pos - 1] < key) { beam[pos] = beam[pos - 1]; --pos; }
                        beam[pos] = key;
                        ++n_b;
                        top[n_t++] = key;                                                // heappush(top_k)
                        if (n_t > a.k) {                                                 // heappop(top_k): (max dist, min id)
                            int e = 0;
                            for (int i = 1; i < n_t; ++i) {
                                const uint32_t di = key_dbits(top[i]), de = key_dbits(top[e]);
                                if (di > de || (di == de && top[i] < top[e])) e = i;
                            }
                            top[e] = top[n_t - 1];
                            --n_t;
                        }
                        wd = 0u;
                        for (int i = 0; i < n_t; ++i) wd = max(wd, key_dbits(top[i]));
                    }
                    state[0] = n_b; state[1] = n_t; state[2] = (int)wd;
                }
                __syncwarp();
            }
            if (lane == 0 && state[0] > a.bw) state[0] = a.bw;             // :595-596: the beam_width WORST entries stay
            __syncwarp();
        }

        // :599-601: top_k as (dist, id) ascending (the reference's order inside an exact-distance tie is heap-layout dependent)
        if (lane == 0) {
            const int n_t = state[1];
            for (int i = 1; i < n_t; ++i) {
                const u64 x = top[i];
                int j = i - 1;
                while (j >= 0 && top[j] > x) { top[j + 1] = top[j]; --j; }
                top[j + 1] = x;
            }
            int o = 0;
            for (int i = 0; i < n_t; ++i) {
                const uint32_t id = key_id(top[i]);
                if (a.deleted && a.deleted[id]) continue;
                const float d = key_dist(top[i]);
                a.out_ids[b * a.k + o] = (int32_t)id;
                if (a.out_dist) a.out_dist[b * a.k + o] = a.sqrt_out ? __fsqrt_rn(d) : d;
                ++o;
            }
            for (; o < a.k; ++o) {
                a.out_ids[b * a.k + o] = -1;
                if (a.out_dist) a.out_dist[b * a.k + o] = __int_as_float(0x7f800000);
            }
            if (a.out_hops) a.out_hops[b] = hops;
            if (a.out_visited) a.out_visited[b] = nvis;
        }
        __syncwarp();
    }
}

// number of resident CTAs (= visited bitmaps) for a batch of B queries, and the bytes of one bitmap
void beam_c_plan(const dr_index *h, int64_t B, long long *grid_out, size_t *bitmap_bytes_out) {
    const size_t per_cta = (size_t)((h->N + 31) / 32) * 4;
    long long grid = 2ll * (h->sms > 0 ? h->sms : 148);
    if (grid > B) grid = B;
    const long long fit = (long long)((256ull << 20) / (per_cta ? per_cta : 1));   // at most 256 MB of bitmaps
    if (grid > fit) grid = fit;
    if (grid < 1) grid = 1;
    *grid_out = grid; *bitmap_bytes_out = per_cta;
}

int launch_beam_c(dr_index *h, const float *d_Q, int64_t B, int k, int bw, int dist, int sqrt_out, const float *d_lut,
                  uint32_t *d_bitmaps, int32_t *ids, float *dists, int32_t *hops, int32_t *visited, cudaStream_t s, int64_t start) {
    if (start < 0) start = h->medoid;
    DR_CHECK(start < h->N, "dr_beam_search_c: start node %lld out of range", (long long)start);
    DR_CHECK(k >= 1 && k <= 1024 && bw >= 0 && bw <= 4096, "dr_beam_search_c: k must be in 1..1024, beam_width in 0..4096");
    DR_CHECK(dist == DR_DIST_PQ || dist == DR_DIST_EXACT, "dr_beam_search_c: dist must be DR_DIST_PQ or DR_DIST_EXACT");
    DR_CHECK(h->N < (1ll << 31), "dr_beam_search_c: ids must fit 31 bits");
    if (dist == DR_DIST_PQ) DR_CHECK(h->d_codes && d_lut && h->M > 0, "dr_beam_search_c: PQ distances need codes and a table");
    else DR_CHECK(h->d_vec, "dr_beam_search_c: exact distances need the vectors");
    if (B == 0) return 0;
    BeamCArgs a;
    memset(&a, 0, sizeof(a));
    a.vec = h->d_vec; a.adj = h->d_adj; a.codes = h->d_codes; a.deg = h->d_deg; a.deleted = h->d_deleted;
    a.Q = d_Q; a.lut = dist == DR_DIST_PQ ? d_lut : nullptr;
    a.N = h->N; a.D = h->D; a.R = h->R; a.M = h->M; a.B = B; a.k = k; a.bw = bw; a.dist = dist; a.sqrt_out = sqrt_out;
    a.start = (uint32_t)start;
    a.out_ids = ids; a.out_dist = dists; a.out_hops = hops; a.out_visited = visited;
    a.words = (h->N + 31) / 32;
    long long grid; size_t per_cta;
    beam_c_plan(h, B, &grid, &per_cta);
    DR_CHECK(d_bitmaps, "dr_beam_search_c: no visited bitmaps");
    a.bitmap = d_bitmaps;
    const size_t smem = ((size_t)(dist == DR_DIST_PQ ? 0 : h->D) * 4 + 15) / 16 * 16 + (size_t)(bw + h->R + 1 + k + 1 + 32) * 8 + 16;
    DR_CHECK(smem <= 48 * 1024, "dr_beam_search_c: D / beam_width / k too large for this compatibility kernel (%zu B of shared memory)", smem);
    beam_c_kernel<<<(unsigned)grid, 32, smem, s>>>(a);
    DR_LAUNCHED();
    return 0;
}
