// build.cu — K5: Vamana index construction on the GPU.
//
// Replaces build_vamana_index_cython + greedy_search_fast_cython + robust_prune_fast_cython
// (cython_utils.pyx:269-492).  The reference inserts points strictly one at a time; here each pass
// (alpha = 1.0, then alpha — cython_utils.pyx:296,310) walks a random permutation in batches
// (prefix doubling up to ~2 % of N, as batched Vamana builders do):
//   1. greedy search of every batch point from the medoid on the current graph snapshot, exact fp32
//      distances (search.cu in DR_DIST_EXACT mode, queries addressed through a row map);
//   2. RobustPrune(alpha, R) of  search list ∪ N(p)  -> new out-row of p         (prune_kernel, mode 0);
//   3. reverse edges: pairs (j, p) for j in N(p) are sorted by (target, source) with our own bitonic sort; each target appends the
//      incoming ids while its row has room, otherwise RobustPrune(N(j) ∪ incoming)  (prune_kernel, mode 1).
// The prune rule is the reference's: candidates sorted by (d, id); after selecting p*, a candidate p'
// is dropped when alpha * d2(p*, p') <= d2(p, p')  (alpha on SQUARED distances, cython_utils.pyx:483).
// Unlike the reference, the search list is a true best-L list (not the FIFO window of :400-423) and the
// stale-tail re-selection of :460-466 is not reproduced; graphs are compared by recall.
#include "common.cuh"

#ifndef DR_PRUNE_X2
#define DR_PRUNE_X2 1   // RobustPrune distance loops: two candidate rows per warp in flight
#endif
#ifndef DR_BUILD_HASH
#define DR_BUILD_HASH 4096   // visited-table slots in shared memory of the build's search (0 = the search's own sizing, 16384 at R = 64);
                             // a query that visits more spills to the search's per-CTA global table
#endif
#ifndef DR_BUILD_THREADS
#define DR_BUILD_THREADS 128 // threads per CTA of the build's search: four small CTAs per SM beat two of 256 threads (1M x 768, R = 64:
                             // 3.17 -> 2.75 s; 64 threads: 2.67-2.96 s)
#endif
#ifndef DR_BUILD_W
#define DR_BUILD_W 4   // list entries expanded per step by the build's batched search
#endif


#include <algorithm>
#include <random>
#include <vector>

#define PR_CMAX 320      // candidate capacity per prune (>= L + R)
#ifndef PR_THREADS
#define PR_THREADS 256   // 8 warps per pruned point (128: 3.57 s, 256: 3.17 s, 384: 3.27 s, 512: 3.45 s for 1M x 768, R = 64)
#endif
#ifndef DR_PRUNE_OCC
#define DR_PRUNE_OCC 0   // > 0: cap on resident prune CTAs per SM (experiment: fewer items in flight = candidate rows stay in L2)
#endif

struct PruneArgs {
    const float *X; int D;
    uint32_t *adj; int32_t *deg; int R; int stride; float alpha;  // stride: adjacency row pitch (== R in the build)
    int mode;  // 0: insert prune (search list ∪ N(p)); 1: reverse edges (N(j) ∪ incoming)
    // mode 0
    const int32_t *nodes; const int32_t *list_ids; const float *list_dist; const int32_t *list_len; int L;
    // mode 1
    const u64 *pairs; const int32_t *heads; const int32_t *n_heads; int n_pairs;   // sorted target << 32 | source keys
    int n_items;
    int *truncated;   // mode 1: targets whose incoming run did not fit PR_CMAX candidates (the rest of the run is dropped); may be NULL
};
std::atomic<long long> g_build_truncated{0};   // of the last dr_vamana_build in this process (dr_vamana_build_last_truncated)

__global__ void __launch_bounds__(PR_THREADS) prune_kernel(const PruneArgs a) {
    extern __shared__ __align__(16) float s_vecs[];  // [D] point p, [D] current p*
    float *s_p = s_vecs, *s_star = s_vecs + a.D;
    __shared__ uint32_t s_id[PR_CMAX];
    __shared__ float s_d[PR_CMAX];
    __shared__ u64 s_key[PR_CMAX];
    __shared__ uint32_t s_sid[PR_CMAX];
    __shared__ float s_sd[PR_CMAX];
    __shared__ unsigned char s_flag[PR_CMAX];  // bit0 known distance, bit1 duplicate; later: alive
    __shared__ uint32_t s_sel[128];
    __shared__ int s_n, s_nu, s_nalive;

    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, wid = tid >> 5, nw = nt >> 5;
    const int D = a.D, R = a.R, RS = a.stride;
    const int n_items = a.mode == 0 ? a.n_items : *a.n_heads;

    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        __syncthreads();
        // ---- gather candidates -----------------------------------------------------------------------
        uint32_t node;
        int n_old = 0;
        if (a.mode == 0) {
            node = (uint32_t)a.nodes[item];
            const int ll = a.list_len[item];
            const int dg = a.deg[node];
            for (int i = tid; i < ll && i < PR_CMAX; i += nt) {
                s_id[i] = (uint32_t)a.list_ids[(size_t)item * a.L + i];
                s_d[i] = a.list_dist[(size_t)item * a.L + i];
                s_flag[i] = 1;
            }
            for (int i = tid; i < dg && ll + i < PR_CMAX; i += nt) {
                s_id[ll + i] = a.adj[(size_t)node * RS + i];
                s_flag[ll + i] = 0;
            }
            if (tid == 0) s_n = min(ll + dg, PR_CMAX);
        } else {
            const int h0 = a.heads[item];
            node = (uint32_t)(a.pairs[h0] >> 32);
            const int dg = a.deg[node];
            n_old = dg;
            for (int i = tid; i < dg; i += nt) { s_id[i] = a.adj[(size_t)node * RS + i]; s_flag[i] = 0; }
            // incoming run: pairs h0.. while the key stays the same
            if (tid == 0) {
                int m = dg;
                int i = h0;
                for (; i < a.n_pairs && (uint32_t)(a.pairs[i] >> 32) == node && m < PR_CMAX; ++i) { s_id[m] = (uint32_t)a.pairs[i]; s_flag[m] = 0; ++m; }
                if (a.truncated && i < a.n_pairs && (uint32_t)(a.pairs[i] >> 32) == node) atomicAdd(a.truncated, 1);   // reverse edges beyond the capacity
                s_n = m;
            }
        }
        __syncthreads();
        const int n = s_n;
        // duplicates (keep the first occurrence) and self
        for (int i = tid; i < n; i += nt) {
            uint32_t id = s_id[i];
            bool dup = (id == node);
            for (int j = 0; j < i && !dup; ++j) dup = (s_id[j] == id);
            if (dup) s_flag[i] |= 2;
        }
        __syncthreads();
        if (a.mode == 1) {
            // room left: append the new ids in order, no distances needed
            int nu = 0;
            for (int i = 0; i < n; ++i) nu += (s_flag[i] & 2) ? 0 : 1;   // n is small; every thread counts (uniform)
            if (nu <= R) {
                if (tid == 0) {
                    int m = n_old;
                    for (int i = n_old; i < n; ++i)
                        if (!(s_flag[i] & 2)) a.adj[(size_t)node * RS + m++] = s_id[i];
                    a.deg[node] = m;
                }
                continue;
            }
        }
        // ---- distances to p ------------------------------------------------------------------------------
        for (int i = tid; i < D; i += nt) s_p[i] = __ldg(a.X + (size_t)node * D + i);
        __syncthreads();
#if DR_PRUNE_X2
        for (int i = wid; i < n;) {   // two unknown candidates per warp at a time (twice the bytes in flight; same sums)
            while (i < n && (s_flag[i] & 3)) i += nw;  // known or duplicate
            if (i >= n) break;
            int i2 = i + nw;
            while (i2 < n && (s_flag[i2] & 3)) i2 += nw;
            if (i2 < n) {
                float dA, dB;
                warp_l2sq_x2(a.X + (size_t)s_id[i] * D, a.X + (size_t)s_id[i2] * D, s_p, D, lane, dA, dB);
                if (lane == 0) { s_d[i] = dA; s_d[i2] = dB; }
            } else {
                const float d = warp_l2sq(a.X + (size_t)s_id[i] * D, s_p, D, lane);
                if (lane == 0) s_d[i] = d;
            }
            i = i2 + nw;
        }
#else
        for (int i = wid; i < n; i += nw) {
            if (s_flag[i] & 3) continue;  // known or duplicate
            float d = warp_l2sq(a.X + (size_t)s_id[i] * D, s_p, D, lane);
            if (lane == 0) s_d[i] = d;
        }
#endif
        __syncthreads();
        // ---- sort by (d, id): rank counting ----------------------------------------------------------------
        for (int i = tid; i < n; i += nt) s_key[i] = (s_flag[i] & 2) ? DR_KEY_MAX : make_key(s_d[i], s_id[i]);
        if (tid == 0) s_nu = 0;
        __syncthreads();
        for (int i = tid; i < n; i += nt) {
            u64 key = s_key[i];
            if (key == DR_KEY_MAX) continue;
            int pos = 0;
            for (int j = 0; j < n; ++j) pos += (s_key[j] < key) ? 1 : 0;
            s_sid[pos] = s_id[i];
            s_sd[pos] = s_d[i];
            atomicAdd(&s_nu, 1);
        }
        __syncthreads();
        const int nu = s_nu;
        for (int i = tid; i < nu; i += nt) s_flag[i] = 1;  // alive
        if (tid == 0) s_nalive = nu;
        __syncthreads();
        // ---- greedy alpha-prune ---------------------------------------------------------------------------
        int cnt = 0;
        for (int i = 0; i < nu && cnt < R; ++i) {
            if (!s_flag[i]) continue;  // uniform: flags only change between barriers
            if (tid == 0) s_sel[cnt] = s_sid[i];
            ++cnt;
            if (cnt == R) break;
            if (s_nalive <= cnt) continue;  // everything still alive is already selected or will be without tests
            for (int t = tid; t < D; t += nt) s_star[t] = __ldg(a.X + (size_t)s_sid[i] * D + t);
            __syncthreads();
#if DR_PRUNE_X2
            for (int j = i + 1 + wid; j < nu;) {   // entries j, j + nw, ... belong to this warp alone: their flags only change here
                while (j < nu && !s_flag[j]) j += nw;
                if (j >= nu) break;
                int j2 = j + nw;
                while (j2 < nu && !s_flag[j2]) j2 += nw;
                if (j2 < nu) {
                    float dA, dB;
                    warp_l2sq_x2(a.X + (size_t)s_sid[j] * D, a.X + (size_t)s_sid[j2] * D, s_star, D, lane, dA, dB);
                    __syncwarp();             // every lane has read s_flag[j], s_flag[j2] before lane 0 clears them
                    if (lane == 0) {
                        int killed = 0;
                        if (__fmul_rn(a.alpha, dA) <= s_sd[j]) { s_flag[j] = 0; ++killed; }
                        if (__fmul_rn(a.alpha, dB) <= s_sd[j2]) { s_flag[j2] = 0; ++killed; }
                        if (killed) atomicSub(&s_nalive, killed);
                    }
                } else {
                    const float d = warp_l2sq(a.X + (size_t)s_sid[j] * D, s_star, D, lane);
                    __syncwarp();
                    if (lane == 0 && __fmul_rn(a.alpha, d) <= s_sd[j]) { s_flag[j] = 0; atomicSub(&s_nalive, 1); }
                }
                j = j2 + nw;
            }
#else
            for (int j = i + 1 + wid; j < nu; j += nw) {
                if (!s_flag[j]) continue;
                float d = warp_l2sq(a.X + (size_t)s_sid[j] * D, s_star, D, lane);
                __syncwarp();
                if (lane == 0 && __fmul_rn(a.alpha, d) <= s_sd[j]) { s_flag[j] = 0; atomicSub(&s_nalive, 1); }
            }
#endif
            __syncthreads();
        }
        __syncthreads();
        for (int t = tid; t < cnt; t += nt) a.adj[(size_t)node * RS + t] = s_sel[t];
        if (tid == 0) a.deg[node] = cnt;
    }
}

// (target, source) pairs of the batch's new out-rows as 64-bit keys target << 32 | source; unused slots and the padding up to the
// sort's power-of-two length get the all-ones sentinel (sorts last)
__global__ void emit_pairs_kernel(const int32_t *__restrict__ nodes, int n_items, const uint32_t *__restrict__ adj,
                                  const int32_t *__restrict__ deg, int R, u64 *__restrict__ pairs, int n_padded) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_padded) return;
    u64 key = DR_KEY_MAX;
    if (t < n_items * R) {
        int item = t / R, j = t - item * R;
        uint32_t node = (uint32_t)nodes[item];
        if (j < deg[node]) key = ((u64)adj[(size_t)node * R + j] << 32) | (u64)node;
    }
    pairs[t] = key;
}

// Bitonic sort of the pair keys, our own two kernels (the reverse edges of a batch are grouped by target; sorting the unique 64-bit
// keys also fixes the order inside a group, so the build is deterministic): compare-exchange distances >= BS_TILE go through global
// memory one launch each, everything below runs inside shared memory in one launch per merge size.
#define BS_TILE 2048
__global__ void __launch_bounds__(BS_TILE / 2) bitonic_smem_kernel(u64 *__restrict__ a, int k_first, int k_last) {
    // sorts / merges inside one tile: for k = k_first .. k_last (doubling), j = min(k, BS_TILE) / 2 .. 1
    __shared__ u64 s[BS_TILE];
    const int base = blockIdx.x * BS_TILE, t = threadIdx.x;
    s[t] = a[base + t]; s[t + BS_TILE / 2] = a[base + t + BS_TILE / 2];
    __syncthreads();
    for (int k = k_first; k <= k_last; k <<= 1) {
        for (int j = (k < BS_TILE ? k : BS_TILE) >> 1; j > 0; j >>= 1) {
            const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));          // lower index of this thread's pair
            const int p = i | j;
            const bool up = (((base + i) & k) == 0);
            const u64 x = s[i], y = s[p];
            if ((x > y) == up) { s[i] = y; s[p] = x; }
            __syncthreads();
        }
    }
    a[base + t] = s[t]; a[base + t + BS_TILE / 2] = s[t + BS_TILE / 2];
}
__global__ void bitonic_global_kernel(u64 *__restrict__ a, int n, int k, int j) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n / 2) return;
    const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
    const int p = i | j;
    const bool up = ((i & k) == 0);
    const u64 x = a[i], y = a[p];
    if ((x > y) == up) { a[i] = y; a[p] = x; }
}
static int bitonic_sort_u64(u64 *d, int n_padded, cudaStream_t s) {      // n_padded: a power of two, >= BS_TILE
    const int tiles = n_padded / BS_TILE;
    bitonic_smem_kernel<<<tiles, BS_TILE / 2, 0, s>>>(d, 2, BS_TILE);     // every tile sorted (alternating directions by position)
    DR_LAUNCHED();
    for (int k = BS_TILE * 2; k <= n_padded; k <<= 1) {
        for (int j = k >> 1; j >= BS_TILE; j >>= 1) {
            bitonic_global_kernel<<<(n_padded / 2 + 255) / 256, 256, 0, s>>>(d, n_padded, k, j);
            DR_LAUNCHED();
        }
        bitonic_smem_kernel<<<tiles, BS_TILE / 2, 0, s>>>(d, k, k);       // the remaining distances of this merge size
        DR_LAUNCHED();
    }
    return 0;
}

// first pair of every target run (the run order the prune kernel sees is immaterial: every run is its own work item)
__global__ void heads_kernel(const u64 *__restrict__ pairs, int n, int32_t *__restrict__ heads, int32_t *__restrict__ n_heads) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const u64 k = pairs[i];
    if (k == DR_KEY_MAX) return;
    if (i == 0 || (uint32_t)(pairs[i - 1] >> 32) != (uint32_t)(k >> 32)) heads[atomicAdd(n_heads, 1)] = i;
}

__global__ void finalize_rows_kernel(uint32_t *__restrict__ adj, const int32_t *__restrict__ deg, long long N, int R) {
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= N * R) return;
    long long i = t / R;
    int j = (int)(t - i * R);
    if (j >= deg[i]) adj[t] = 0u;  // DiskANNPersist.save_index pads short rows with 0 (diskann_persist.py:23)
}

struct BuildBufs {
    int32_t *sigma = nullptr, *list_ids = nullptr, *list_len = nullptr, *heads = nullptr, *n_heads = nullptr, *topk = nullptr;
    float *list_dist = nullptr;
    u64 *pairs = nullptr;
    ~BuildBufs() {
        cudaFree(sigma); cudaFree(list_ids); cudaFree(list_len); cudaFree(heads); cudaFree(n_heads); cudaFree(topk);
        cudaFree(list_dist); cudaFree(pairs);
    }
};

int launch_vamana_build(const float *d_X, int64_t N, int D, int R, int L, float alpha, int64_t medoid, uint64_t seed,
                        uint32_t *d_adj, int32_t *d_deg, int device, cudaStream_t s) {
    DR_CHECK(N >= 1 && D >= 1, "dr_vamana_build: bad shape");
    DR_CHECK(R >= 1 && R <= 128, "dr_vamana_build: R must be in 1..128 (got %d)", R);
    DR_CHECK(L >= 1 && L <= 512 && L + R <= PR_CMAX, "dr_vamana_build: need L + R <= %d (L=%d R=%d)", PR_CMAX, L, R);
    DR_CHECK(medoid >= 0 && medoid < N, "dr_vamana_build: medoid out of range");
    DR_CHECK(N < (1ll << 31), "dr_vamana_build: N must be < 2^31");

    // a temporary index view over the graph under construction (owns only its scratch)
    dr_index g;
    g.device = device; g.N = N; g.D = D; g.R = R; g.M = 0; g.medoid = medoid; g.owns = false;
    g.d_vec = const_cast<float *>(d_X); g.d_adj = d_adj; g.d_deg = d_deg;
    cudaDeviceProp prop;
    DR_CUDA(cudaGetDeviceProperties(&prop, device));
    g.sms = prop.multiProcessorCount; g.smem_optin = (int)prop.sharedMemPerBlockOptin;
    struct Guard { dr_index *g; ~Guard() { cudaFree(g->d_counter); cudaFree(g->d_ovf); cudaFree(g->d_lut); cudaFree(g->d_io); } } guard{&g};

    int64_t maxb = N / 50;
    if (maxb > 16384) maxb = 16384;
    if (maxb < 1) maxb = 1;

    BuildBufs b;
    DR_CUDA(cudaMalloc(&b.sigma, (size_t)N * 4));
    DR_CUDA(cudaMalloc(&b.list_ids, (size_t)maxb * L * 4));
    DR_CUDA(cudaMalloc(&b.list_dist, (size_t)maxb * L * 4));
    DR_CUDA(cudaMalloc(&b.list_len, (size_t)maxb * 4));
    DR_CUDA(cudaMalloc(&b.topk, (size_t)maxb * 4));
    const int max_pairs = (int)(maxb * R);
    int max_padded = BS_TILE;
    while (max_padded < max_pairs) max_padded <<= 1;
    DR_CUDA(cudaMalloc(&b.pairs, (size_t)max_padded * 8));
    DR_CUDA(cudaMalloc(&b.heads, (size_t)max_pairs * 4));
    DR_CUDA(cudaMalloc(&b.n_heads, 8));      // [0] head count of the batch, [1] truncated-run counter of the whole build
    DR_CUDA(cudaMemsetAsync(b.n_heads, 0, 8, s));
    DR_CUDA(cudaMemsetAsync(d_deg, 0, (size_t)N * 4, s));
    DR_CUDA(cudaMemsetAsync(d_adj, 0, (size_t)N * R * 4, s));

    dr_search_params sp;
    memset(&sp, 0, sizeof(sp));
    sp.k = 1; sp.L = L; sp.W = DR_BUILD_W; sp.dist = DR_DIST_EXACT; sp.rerank = 0;
    sp.hash_cap = DR_BUILD_HASH; sp.threads = DR_BUILD_THREADS;

    const size_t prune_smem = (size_t)2 * D * 4;
    DR_CUDA(cudaFuncSetAttribute(prune_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)prune_smem));
    int pocc = 0;
    DR_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&pocc, prune_kernel, PR_THREADS, prune_smem));
    DR_CHECK(pocc >= 1, "dr_vamana_build: D=%d too large for the prune kernel", D);
    const int prune_grid_max = g.sms * ((DR_PRUNE_OCC > 0 && DR_PRUNE_OCC < pocc) ? DR_PRUNE_OCC : pocc);

    std::mt19937_64 rng(seed);
    std::vector<int32_t> sigma(N);
    for (int pass = 0; pass < 2; ++pass) {
        for (int64_t i = 0; i < N; ++i) sigma[i] = (int32_t)i;
        std::shuffle(sigma.begin(), sigma.end(), rng);
        DR_CUDA(cudaMemcpyAsync(b.sigma, sigma.data(), (size_t)N * 4, cudaMemcpyHostToDevice, s));
        DR_CUDA(cudaStreamSynchronize(s));  // sigma is reused by the next pass
        const float a_pass = pass == 0 ? 1.0f : alpha;
        int64_t start = 0;
        while (start < N) {
            int64_t bs = (pass == 0) ? std::max<int64_t>(1, std::min<int64_t>(start, maxb)) : maxb;
            if (bs > N - start) bs = N - start;
            const int32_t *nodes = b.sigma + start;
            // 1. search
            if (launch_search(&g, d_X, bs, &sp, nullptr, b.topk, nullptr, nullptr, nullptr, b.list_ids, b.list_dist, b.list_len,
                              nullptr, 0, nullptr, s, nodes)) return 1;
            // 2. prune the batch points
            PruneArgs pa;
            memset(&pa, 0, sizeof(pa));
            pa.X = d_X; pa.D = D; pa.adj = d_adj; pa.deg = d_deg; pa.R = R; pa.stride = R; pa.alpha = a_pass;
            pa.mode = 0; pa.nodes = nodes; pa.list_ids = b.list_ids; pa.list_dist = b.list_dist; pa.list_len = b.list_len; pa.L = L;
            pa.n_items = (int)bs;
            prune_kernel<<<(int)std::min<int64_t>(bs, prune_grid_max), PR_THREADS, prune_smem, s>>>(pa);
            DR_LAUNCHED();
            // 3. reverse edges
            const int np = (int)(bs * R);
            int npad = BS_TILE;
            while (npad < np) npad <<= 1;
            emit_pairs_kernel<<<(npad + 255) / 256, 256, 0, s>>>(nodes, (int)bs, d_adj, d_deg, R, b.pairs, npad);
            DR_LAUNCHED();
            if (bitonic_sort_u64(b.pairs, npad, s)) return 1;
            DR_CUDA(cudaMemsetAsync(b.n_heads, 0, 4, s));
            heads_kernel<<<(np + 255) / 256, 256, 0, s>>>(b.pairs, np, b.heads, b.n_heads);
            DR_LAUNCHED();
            PruneArgs pr = pa;
            pr.mode = 1; pr.pairs = b.pairs; pr.heads = b.heads; pr.n_heads = b.n_heads; pr.n_pairs = np;
            pr.truncated = b.n_heads + 1;
            prune_kernel<<<(int)std::min<int64_t>(np, prune_grid_max), PR_THREADS, prune_smem, s>>>(pr);
            DR_LAUNCHED();
            start += bs;
        }
    }
    const long long tot = (long long)N * R;
    finalize_rows_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, s>>>(d_adj, d_deg, N, R);
    DR_LAUNCHED();
    int trunc = 0;
    DR_CUDA(cudaMemcpyAsync(&trunc, b.n_heads + 1, 4, cudaMemcpyDeviceToHost, s));
    DR_CUDA(cudaStreamSynchronize(s));
    g_build_truncated.store(trunc);
    return 0;
}

// RobustPrune of one point against an explicit candidate set (robust_prune_cython, cython_utils.pyx:124-167, and
// the cdef version :435-492).  d_X holds the n candidate rows followed by the point itself (row n); candidate
// "ids" inside the kernel are the local row numbers, so the caller passes candidates in ascending real-id order
// to keep the (distance, id) tie order.  d_sel receives the selected local rows, *d_cnt their number.
int launch_prune_one(const float *d_X, int n, int D, float alpha, int R, uint32_t *d_row, int32_t *d_deg, cudaStream_t s) {
    DR_CHECK(n >= 0 && n <= PR_CMAX, "dr_robust_prune: at most %d candidates (got %d)", PR_CMAX, n);
    DR_CHECK(R >= 1 && R <= 128, "dr_robust_prune: R must be in 1..128");
    // a 1-node "graph": node id n, whose current row lists the candidates 0..n-1
    PruneArgs pa;
    memset(&pa, 0, sizeof(pa));
    pa.X = d_X; pa.D = D; pa.adj = d_row - (size_t)n * (n > R ? n : R); pa.deg = d_deg - n;  // row/deg of node n land on d_row/d_deg
    pa.R = R; pa.stride = (n > R ? n : R); pa.alpha = alpha; pa.mode = 0;
    // scratch of this call, on the current device: [0..1] = an empty search list (ids / length 0), [2] = the node id n
    struct Scratch { int32_t *p = nullptr; ~Scratch() { if (p) cudaFree(p); } } sc;
    DR_CUDA(cudaMalloc(&sc.p, 16));
    const int32_t init[4] = {0, 0, n, 0};
    DR_CUDA(cudaMemcpyAsync(sc.p, init, 16, cudaMemcpyHostToDevice, s));
    pa.nodes = sc.p + 2; pa.list_ids = sc.p; pa.list_dist = (const float *)sc.p; pa.list_len = sc.p; pa.L = 1; pa.n_items = 1;
    const size_t smem = (size_t)2 * D * 4;
    DR_CUDA(cudaFuncSetAttribute(prune_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    prune_kernel<<<1, PR_THREADS, smem, s>>>(pa);
    DR_LAUNCHED();
    DR_CUDA(cudaStreamSynchronize(s));
    return 0;
}
