// distance.cu — K4 fused row-wise distance kernels and K6 the cross-shard top-k merge.
//
// K4 replaces l2_distance_fast_cython (cython_utils.pyx:18-24), cosine_similarity_cython (:53-70, a
// distance: 1 - cos, 0 when a norm is 0), pq_distance_fast_cython (:26-51) and the dot product, over n
// row pairs in one launch.  One warp per pair, coalesced float4 loads, butterfly reduction; HBM-bound.
#include "common.cuh"

// op: 0 = squared L2, 1 = dot, 2 = cosine distance
__global__ void __launch_bounds__(256) rowdist_kernel(const float *__restrict__ A, const float *__restrict__ Bm, long long n,
                                                      long long nb, int D, int op, float *__restrict__ out) {
    const int lane = threadIdx.x & 31;
    const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (w >= n) return;
    const float *a = A + (size_t)w * D;
    const float *b = Bm + (size_t)(nb == 1 ? 0 : w) * D;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f;
    if ((D & 3) == 0) {
#pragma unroll 4
        for (int base = lane * 4; base < D; base += 128) {
            float4 x = ldg_f4(a + base), y = ldg_f4(b + base);
            float xs[4] = {x.x, x.y, x.z, x.w}, ys[4] = {y.x, y.y, y.z, y.w};
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                if (op == 0) { float d = __fsub_rn(xs[c], ys[c]); s0 = __fmaf_rn(d, d, s0); }
                else {
                    s0 = __fmaf_rn(xs[c], ys[c], s0);
                    if (op == 2) { s1 = __fmaf_rn(xs[c], xs[c], s1); s2 = __fmaf_rn(ys[c], ys[c], s2); }
                }
            }
        }
    } else {
        for (int i = lane; i < D; i += 32) {
            float x = __ldg(a + i), y = __ldg(b + i);
            if (op == 0) { float d = __fsub_rn(x, y); s0 = __fmaf_rn(d, d, s0); }
            else {
                s0 = __fmaf_rn(x, y, s0);
                if (op == 2) { s1 = __fmaf_rn(x, x, s1); s2 = __fmaf_rn(y, y, s2); }
            }
        }
    }
    s0 = warp_sum_butterfly(s0);
    if (op == 2) { s1 = warp_sum_butterfly(s1); s2 = warp_sum_butterfly(s2); }
    if (lane == 0) {
        float r = s0;
        if (op == 2) {
            if (s1 == 0.0f || s2 == 0.0f) r = 0.0f;
            else r = (float)(1.0 - ((double)s0 / (sqrt((double)s1) * sqrt((double)s2))));
        }
        out[w] = r;
    }
}

int launch_rowdist(const float *dA, const float *dB, int64_t n, int64_t nb, int D, int op, float *d_out, cudaStream_t s) {
    if (n == 0) return 0;
    const long long threads = (long long)n * 32;
    rowdist_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, s>>>(dA, dB, n, nb, D, op, d_out);
    DR_LAUNCHED();
    return 0;
}

// symmetric PQ distance: one thread per pair, fp32 sequential like the reference
__global__ void sdc_kernel(const float *__restrict__ codebook, const uint8_t *__restrict__ c1, const uint8_t *__restrict__ c2,
                           long long n, int M, int ds, float *__restrict__ out) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float total = 0.0f;
    for (int m = 0; m < M; ++m) {
        const float *a = codebook + ((size_t)m * 256 + c1[(size_t)i * M + m]) * ds;
        const float *b = codebook + ((size_t)m * 256 + c2[(size_t)i * M + m]) * ds;
        float sub = 0.0f;
        for (int j = 0; j < ds; ++j) {
            float d = __fsub_rn(__ldg(a + j), __ldg(b + j));
            sub = __fadd_rn(sub, __fmul_rn(d, d));
        }
        total = __fadd_rn(total, sub);
    }
    out[i] = total;
}

int launch_sdc(const float *d_codebook, const uint8_t *d_c1, const uint8_t *d_c2, int64_t n, int M, int ds, float *d_out,
               cudaStream_t s) {
    if (n == 0) return 0;
    sdc_kernel<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(d_codebook, d_c1, d_c2, n, M, ds, d_out);
    DR_LAUNCHED();
    return 0;
}

// K6: k-way merge of G per-shard ascending top-k lists per query: one warp per query, G*k <= 1024.
// Rank by (dist, id); empty slots (id < 0) sort last.
__global__ void __launch_bounds__(128) topk_merge_kernel(const int32_t *__restrict__ ids, const float *__restrict__ dist, int G,
                                                         long long B, int k, int32_t *__restrict__ out_ids,
                                                         float *__restrict__ out_dist) {
    extern __shared__ u64 s_keys[];  // [warps][G*k]
    const int lane = threadIdx.x & 31, wl = threadIdx.x >> 5;
    const long long q = (long long)blockIdx.x * (blockDim.x >> 5) + wl;
    if (q >= B) return;
    const int n = G * k;
    u64 *keys = s_keys + (size_t)wl * n;
    for (int i = lane; i < n; i += 32) {
        int g = i / k, j = i - g * k;
        int32_t id = ids[((size_t)g * B + q) * k + j];
        float d = dist[((size_t)g * B + q) * k + j];
        keys[i] = id < 0 ? DR_KEY_MAX : (((u64)f2ord(d + 0.0f) << 32) | (uint32_t)id);
    }
    __syncwarp();
    for (int i = lane; i < n; i += 32) {
        u64 key = keys[i];
        int pos = 0;
        for (int j = 0; j < n; ++j) pos += (keys[j] < key || (keys[j] == key && j < i)) ? 1 : 0;
        if (pos < k) {
            bool empty = key == DR_KEY_MAX;
            out_ids[(size_t)q * k + pos] = empty ? -1 : (int32_t)(key & 0xFFFFFFFFull);
            out_dist[(size_t)q * k + pos] = empty ? __int_as_float(0x7f800000) : ord2f((uint32_t)(key >> 32));
        }
    }
}

int launch_topk_merge(const int32_t *d_ids, const float *d_dist, int G, int64_t B, int k, int32_t *d_out_ids,
                      float *d_out_dist, cudaStream_t s) {
    DR_CHECK(G >= 1 && k >= 1 && G * k <= 1024, "dr_topk_merge: need G*k <= 1024 (G=%d k=%d)", G, k);
    if (B == 0) return 0;
    const int warps = 4;
    size_t smem = (size_t)warps * G * k * 8;
    topk_merge_kernel<<<(unsigned)((B + warps - 1) / warps), warps * 32, smem, s>>>(d_ids, d_dist, G, B, k, d_out_ids, d_out_dist);
    DR_LAUNCHED();
    return 0;
}

// Index-sharded exchange, packed: one 64-bit key per (query, rank-in-list) = f2ord(dist) << 32 | global id; empty = all ones.
// Send layout [G][Bq][k]: block g holds the lists of the queries rank g reduces (contiguous slices of the batch, the first B % G
// ranks own one query more), rows past the slice are empty.  One all-to-all of this buffer replaces two (ids, distances).
__global__ void topk_pack_kernel(const int32_t *__restrict__ ids, const float *__restrict__ dist, long long B, int k, long long id_offset,
                                 int G, long long Bq, u64 *__restrict__ out) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)G * Bq * k) return;
    const int j = (int)(t % k);
    const long long r = (t / k) % Bq;
    const int g = (int)(t / ((long long)k * Bq));
    const long long per = B / G, extra = B % G;
    const long long lo = g * per + (g < extra ? g : extra), cnt = per + (g < extra ? 1 : 0);
    u64 key = DR_KEY_MAX;
    if (r < cnt) {
        const int32_t id = ids[(size_t)(lo + r) * k + j];
        if (id >= 0) key = ((u64)f2ord(dist[(size_t)(lo + r) * k + j] + 0.0f) << 32) | (u64)(uint32_t)(id + id_offset);
    }
    out[t] = key;
}
// k-way merge of G packed lists per query, keys [G][Bq][k] -> out [Bq][k]; one warp per query, G*k <= 1024
__global__ void __launch_bounds__(128) topk_merge_keys_kernel(const u64 *__restrict__ keys_in, int G, long long Bq, int k,
                                                              int32_t *__restrict__ out_ids, float *__restrict__ out_dist) {
    extern __shared__ u64 s_keys[];
    const int lane = threadIdx.x & 31, wl = threadIdx.x >> 5;
    const long long q = (long long)blockIdx.x * (blockDim.x >> 5) + wl;
    if (q >= Bq) return;
    const int n = G * k;
    u64 *keys = s_keys + (size_t)wl * n;
    for (int i = lane; i < n; i += 32) {
        const int g = i / k, j = i - g * k;
        keys[i] = keys_in[((size_t)g * Bq + q) * k + j];
    }
    __syncwarp();
    for (int i = lane; i < n; i += 32) {
        const u64 key = keys[i];
        int pos = 0;
        for (int j = 0; j < n; ++j) pos += (keys[j] < key || (keys[j] == key && j < i)) ? 1 : 0;
        if (pos < k) {
            const bool empty = key == DR_KEY_MAX;
            out_ids[(size_t)q * k + pos] = empty ? -1 : (int32_t)(key & 0xFFFFFFFFull);
            out_dist[(size_t)q * k + pos] = empty ? __int_as_float(0x7f800000) : ord2f((uint32_t)(key >> 32));
        }
    }
}
int launch_topk_pack(const int32_t *d_ids, const float *d_dist, int64_t B, int k, int64_t id_offset, int G, u64 *d_out, cudaStream_t s) {
    DR_CHECK(G >= 1 && k >= 1, "dr_topk_pack: bad G / k");
    const long long Bq = (B + G - 1) / G, tot = (long long)G * Bq * k;
    if (tot == 0) return 0;
    topk_pack_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, s>>>(d_ids, d_dist, B, k, id_offset, G, Bq, d_out);
    DR_LAUNCHED();
    return 0;
}
int launch_topk_merge_keys(const u64 *d_keys, int G, int64_t Bq, int k, int32_t *d_out_ids, float *d_out_dist, cudaStream_t s) {
    DR_CHECK(G >= 1 && k >= 1 && G * k <= 1024, "dr_topk_merge_keys: need G*k <= 1024 (G=%d k=%d)", G, k);
    if (Bq == 0) return 0;
    const int warps = 4;
    topk_merge_keys_kernel<<<(unsigned)((Bq + warps - 1) / warps), warps * 32, (size_t)warps * G * k * 8, s>>>(d_keys, G, Bq, k, d_out_ids,
                                                                                                               d_out_dist);
    DR_LAUNCHED();
    return 0;
}

// medoid (cython_utils.pyx:210-263): sums[s] = sum_j ||x_sample_s - x_j||  (fp32 inner, fp64 outer).
// grid (ns, chunks over N); one warp per (sample, point) pair inside, double atomics per CTA.
__global__ void __launch_bounds__(256) medoid_kernel(const float *__restrict__ X, long long N, int D,
                                                     const int32_t *__restrict__ samples, int skip_self,
                                                     double *__restrict__ sums) {
    extern __shared__ float s_x[];  // sample vector
    const int sidx = blockIdx.x;
    const long long srow = samples[sidx];
    for (int i = threadIdx.x; i < D; i += blockDim.x) s_x[i] = __ldg(X + (size_t)srow * D + i);
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    double acc = 0.0;
    for (long long j = (long long)blockIdx.y * nw + wid; j < N; j += (long long)gridDim.y * nw) {
        if (skip_self && j == srow) continue;
        float d2 = warp_l2sq(X + (size_t)j * D, s_x, D, lane);
        acc += sqrt((double)d2);
    }
    __shared__ double s_acc[8];
    if (lane == 0) s_acc[wid] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < nw; ++w) t += s_acc[w];
        atomicAdd(&sums[sidx], t);
    }
}

int launch_medoid(const float *d_X, int64_t N, int D, const int32_t *d_samples, int ns, int skip_self, double *d_sums,
                  cudaStream_t s) {
    DR_CUDA(cudaMemsetAsync(d_sums, 0, sizeof(double) * ns, s));
    int chunks = (int)((N + 8 * 64 - 1) / (8 * 64));
    if (chunks > 64) chunks = 64;
    if (chunks < 1) chunks = 1;
    dim3 grid(ns, chunks);
    medoid_kernel<<<grid, 256, (size_t)D * 4, s>>>(d_X, N, D, d_samples, skip_self, d_sums);
    DR_LAUNCHED();
    return 0;
}
