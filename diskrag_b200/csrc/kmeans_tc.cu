// kmeans_tc.cu — K3 on the tensor cores: the assignment step of the per-subspace k-means that trains the PQ
// codebooks (DiskANNPQ.fit -> sklearn KMeans.fit, pydiskann/pq/fast_pq.py:197-243), with tcgen05.mma (TF32).
//
// For subspace m the step is one dense contraction, [rows x ds] . [ds x 256]: D[r][c] = ||C_c||^2 - 2 x_r . C_c
// (the ||x_r||^2 term does not change the argmin).  One CTA owns one subspace and a run of 128-row tiles: the 256
// centroids (B operand, plus the extra K-step that adds ||C_c||^2 as an exact two-term TF32 split) are staged in shared
// memory ONCE, every tile stages its 128 x ds rows (A = -2 x, plus [1, 1, 0..]), one N = 256 MMA per K-step fills 256 TMEM
// columns, and two threads per row (TMEM lane r, columns 0..127 / 128..255) take the argmin with tcgen05.ld as a packed
// (distance, centroid) integer min; the rows of the next tile are prefetched while this one is reduced.  The Lloyd update is unchanged:
// exact fp32 sums of the rows per winning centroid (shared-memory atomics, one flush per CTA), pq.cu:update_kernel.
// TF32 products can flip an assignment between two near-equidistant centroids; k-means is judged by quantisation MSE
// (tests/test_kernels_gpu.py, tests/test_lut_tc_gpu.py), and the final ENCODE stays on the exact fp32 path (pq.cu).
#include "tc_common.cuh"

#define KM_ROWS 128
#define KM_THREADS 256   // two threads per row: warps 0-3 take TMEM columns 0..127 of their 32-lane quarter, warps 4-7 columns 128..255
#define KM_TILES 16      // row tiles per CTA: one flush of the partial sums per 2048 rows

namespace {

struct KmTcArgs {
    const float *X; long long N; int D, M, ds; long long stride;
    const float *codebook; float *sums; int *counts; double *sse;
};

// Epilogue arithmetic: D[r][c] = ||x_r - C_c||^2 >= 0 (the extra K-step adds ||C_c||^2 AND ||x_r||^2 as exact two-term TF32 splits),
// so the fp32 bit pattern orders like a signed integer and (bits & ~0xFF) | c is a packed (distance, centroid) key: one LOP3 + one
// integer min per element instead of compare + two selects, and the lowest index wins among equal (truncated) distances.
// dynamic shared memory: A (nks+1) x 4 KB | B (nks+1) x 8 KB | sums 256 x ds floats | counts 256 ints | keys 2 x 128 ints
template <int DS>   // compile-time sub-dimension (a multiple of 8, <= 64): the row buffers are DS / 4 float4 registers
__global__ void __launch_bounds__(KM_THREADS, 2) kmeans_assign_tc_kernel(const KmTcArgs a) {
    extern __shared__ __align__(1024) unsigned char km_smem[];
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ uint32_t s_tmem;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    constexpr int ds = DS, nks = ds >> 3, nks1 = nks + 1, KC = DS / 4;
    const int D = a.D, m = blockIdx.y;
    const int half = wid >> 2;                 // which 128 columns this thread reduces
    const int row = tid & (KM_ROWS - 1);       // TMEM lane = row of the tile
    const bool owner = tid < KM_ROWS;          // owners stage the A operand of their row and apply the Lloyd update for it
    unsigned char *sA = km_smem;
    unsigned char *sB = sA + nks1 * (KM_ROWS * 32);
    float *s_sum = reinterpret_cast<float *>(sB + nks1 * (256 * 32));
    int *s_cnt = reinterpret_cast<int *>(s_sum + 256 * ds);
    int *s_key = s_cnt + 256;                  // [2][128]
    const uint32_t a_base = smem_u32(sA), b_base = smem_u32(sB);

    if (tid == 0) { mbar_init(&s_bar, 1); fence_mbar_init(); }
    if (wid == 0) tmem_alloc(&s_tmem, 256);
    {   // B operand: the subspace's 256 centroids, one per thread; extra K-step [cn_hi, cn_lo, 1, 1 | 0 0 0 0]
        const int r = tid;
        const float *src = a.codebook + ((size_t)m * 256 + r) * ds;
        float cn = 0.0f;
        for (int kc = 0; kc < (ds >> 2); ++kc) {
            const float4 v = ldg_f4(src + kc * 4);
            cn = __fmaf_rn(v.x, v.x, cn); cn = __fmaf_rn(v.y, v.y, cn); cn = __fmaf_rn(v.z, v.z, cn); cn = __fmaf_rn(v.w, v.w, cn);
            *reinterpret_cast<float4 *>(sB + (kc >> 1) * (256 * 32) + core_off(r, kc & 1)) = tf32_rna4(v);
        }
        const float ch = tf32_hi(cn);
        unsigned char *x = sB + nks * (256 * 32);
        *reinterpret_cast<float4 *>(x + core_off(r, 0)) = make_float4(ch, cn - ch, 1.0f, 1.0f);
        *reinterpret_cast<float4 *>(x + core_off(r, 1)) = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    }
    for (int i = tid; i < 256 * ds; i += KM_THREADS) s_sum[i] = 0.0f;
    s_cnt[tid] = 0;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem;
    const uint32_t tcol = tmem + ((uint32_t)((wid & 3) * 32) << 16) + (uint32_t)(half * 128);
    uint32_t phase = 0;
    float sse_local = 0.0f;

    const long long tile0 = (long long)blockIdx.x * KM_TILES;
    // software pipeline over the tiles: the rows of tile t + 1 travel from HBM while tile t is reduced
    float4 xr[KC];                             // this owner's row of the tile in flight
    bool live = false;
    auto fetch = [&](int t) {
        const long long i = (tile0 + t) * KM_ROWS + row;
        live = i < a.N;
        if (owner && live) {
            const float *x = a.X + (size_t)(i * a.stride) * D + (size_t)m * ds;
#pragma unroll
            for (int kc = 0; kc < KC; ++kc) xr[kc] = ldg_f4(x + kc * 4);
        }
    };
    fetch(0);
    for (int t = 0; t < KM_TILES; ++t) {
        const long long r0 = (tile0 + t) * KM_ROWS;
        if (r0 >= a.N) break;                                   // uniform
        const bool live_t = live;
        float4 xc[KC];                                          // the row being reduced (kept for the update)
        if (owner) {
            float xn = 0.0f;
#pragma unroll
            for (int kc = 0; kc < KC; ++kc) {
                {
                    float4 v = live_t ? xr[kc] : make_float4(0.f, 0.f, 0.f, 0.f);
                    xc[kc] = v;
                    xn = __fmaf_rn(v.x, v.x, xn); xn = __fmaf_rn(v.y, v.y, xn); xn = __fmaf_rn(v.z, v.z, xn); xn = __fmaf_rn(v.w, v.w, xn);
                    v.x *= -2.0f; v.y *= -2.0f; v.z *= -2.0f; v.w *= -2.0f;
                    *reinterpret_cast<float4 *>(sA + (kc >> 1) * (KM_ROWS * 32) + core_off(row, kc & 1)) = tf32_rna4(v);
                }
            }
            const float xh = tf32_hi(xn);
            unsigned char *e = sA + nks * (KM_ROWS * 32);
            *reinterpret_cast<float4 *>(e + core_off(row, 0)) = make_float4(1.0f, 1.0f, xh, xn - xh);
            *reinterpret_cast<float4 *>(e + core_off(row, 1)) = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        }
        if (t + 1 < KM_TILES && (tile0 + t + 1) * KM_ROWS < a.N) fetch(t + 1);   // next tile's rows: in flight during this tile
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
            constexpr uint32_t idesc = umma_idesc_tf32(KM_ROWS, 256);
            for (int ks = 0; ks < nks1; ++ks)
                umma_tf32(tmem, umma_desc(a_base + (uint32_t)(ks * (KM_ROWS * 32))), umma_desc(b_base + (uint32_t)(ks * (256 * 32))), idesc,
                          ks > 0 ? 1u : 0u);
            umma_commit(&s_bar);
        }
        mbar_wait(&s_bar, phase);
        phase ^= 1u;
        tc_fence_after();
        int best = 0x7fffffff;
#pragma unroll
        for (int ch = 0; ch < 8; ch += 2) {                     // two 16-column loads in flight per wait
            float v0[16], v1[16];
            tmem_ld16(tcol + (uint32_t)(ch * 16), v0);
            tmem_ld16(tcol + (uint32_t)(ch * 16 + 16), v1);
            tmem_ld_wait();
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                best = min(best, (__float_as_int(v0[q]) & (int)0xFFFFFF00) | (ch * 16 + q));
                best = min(best, (__float_as_int(v1[q]) & (int)0xFFFFFF00) | (ch * 16 + 16 + q));
            }
        }
        s_key[half * KM_ROWS + row] = best | (half << 7);
        tc_fence_before();                                      // TMEM reads of this tile are ordered before the next tile's MMA
        __syncthreads();
        if (owner && live_t) {
            const int k0 = s_key[row], k1 = s_key[KM_ROWS + row];
            const int kb = min(k0, k1);
            const int bi = kb & 0xFF;
#pragma unroll
            for (int kc = 0; kc < KC; ++kc) {
                {
                    const float4 v = xc[kc];
                    atomicAdd(&s_sum[bi * ds + kc * 4 + 0], v.x); atomicAdd(&s_sum[bi * ds + kc * 4 + 1], v.y);
                    atomicAdd(&s_sum[bi * ds + kc * 4 + 2], v.z); atomicAdd(&s_sum[bi * ds + kc * 4 + 3], v.w);
                }
            }
            atomicAdd(&s_cnt[bi], 1);
            sse_local += fmaxf(__int_as_float(kb & (int)0xFFFFFF00), 0.0f);
        }
    }
    tc_fence_before();
    __syncthreads();
    for (int i = tid; i < 256 * ds; i += KM_THREADS)
        if (s_sum[i] != 0.0f) atomicAdd(&a.sums[(size_t)m * 256 * ds + i], s_sum[i]);
    if (s_cnt[tid]) atomicAdd(&a.counts[m * 256 + tid], s_cnt[tid]);
    if (a.sse) {
        const float v = warp_sum_butterfly(sse_local);
        if (lane == 0 && owner) atomicAdd(a.sse, (double)v);
    }
    if (wid == 0) tmem_dealloc(tmem, 256);
}

}  // namespace

// one Lloyd assignment pass over rows i * stride, i < N (accumulates sums / counts / sse like pq.cu:assign_kernel<true>)
int launch_kmeans_assign_tc(const float *d_X, long long N, int D, int M, long long stride, const float *d_codebook, float *d_sums,
                            int *d_counts, double *d_sse, cudaStream_t s) {
    const int ds = D / M, nks1 = ds / 8 + 1;
    DR_CHECK((ds & 7) == 0, "k-means on tensor cores needs a sub-dimension that is a multiple of 8 (got %d)", ds);
    const int smem = nks1 * (KM_ROWS * 32) + nks1 * (256 * 32) + 256 * ds * 4 + 256 * 4 + 2 * KM_ROWS * 4;
    DR_CHECK(smem <= 100 * 1024, "k-means on tensor cores: sub-dimension %d too large", ds);
    void (*kern)(const KmTcArgs) = nullptr;
    switch (ds) {
        case 8: kern = kmeans_assign_tc_kernel<8>; break;
        case 16: kern = kmeans_assign_tc_kernel<16>; break;
        case 24: kern = kmeans_assign_tc_kernel<24>; break;
        case 32: kern = kmeans_assign_tc_kernel<32>; break;
        case 40: kern = kmeans_assign_tc_kernel<40>; break;
        case 48: kern = kmeans_assign_tc_kernel<48>; break;
        case 56: kern = kmeans_assign_tc_kernel<56>; break;
        case 64: kern = kmeans_assign_tc_kernel<64>; break;
        default: break;
    }
    DR_CHECK(kern, "k-means on tensor cores: sub-dimension %d not instantiated (multiples of 8 up to 64)", ds);
    DR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    KmTcArgs a;
    a.X = d_X; a.N = N; a.D = D; a.M = M; a.ds = ds; a.stride = stride; a.codebook = d_codebook; a.sums = d_sums; a.counts = d_counts;
    a.sse = d_sse;
    const long long tiles = (N + KM_ROWS - 1) / KM_ROWS;
    dim3 grid((unsigned)((tiles + KM_TILES - 1) / KM_TILES), (unsigned)M);
    kern<<<grid, KM_THREADS, smem, s>>>(a);
    DR_LAUNCHED();
    return 0;
}
