// kmeans_tc.cu — K3 on the tensor cores: the assignment step of the per-subspace k-means that trains the PQ
// codebooks (DiskANNPQ.fit -> sklearn KMeans.fit, pydiskann/pq/fast_pq.py:197-243), with tcgen05.mma (TF32).
//
// For subspace m the step is one dense contraction, [rows x ds] . [ds x 256]: D[r][c] = ||C_c||^2 - 2 x_r . C_c
// (the ||x_r||^2 term does not change the argmin).  One CTA owns one subspace and a run of 128-row tiles: the 256
// centroids (B operand, plus the extra K-step that adds ||C_c||^2 as an exact two-term TF32 split) are staged in shared
// memory ONCE, every tile stages its 128 x ds rows (A = -2 x, plus [1, 1, 0..]), one N = 256 MMA per K-step fills 256 TMEM
// columns, and thread r (= TMEM lane r) takes the argmin of its row with tcgen05.ld.  The Lloyd update is unchanged:
// exact fp32 sums of the rows per winning centroid (shared-memory atomics, one flush per CTA), pq.cu:update_kernel.
// TF32 products can flip an assignment between two near-equidistant centroids; k-means is judged by quantisation MSE
// (tests/test_kernels_gpu.py, tests/test_lut_tc_gpu.py), and the final ENCODE stays on the exact fp32 path (pq.cu).
#include "tc_common.cuh"

#define KM_ROWS 128
#define KM_TILES 16      // row tiles per CTA: one flush of the partial sums per 2048 rows

namespace {

struct KmTcArgs {
    const float *X; long long N; int D, M, ds; long long stride;
    const float *codebook; float *sums; int *counts; double *sse;
};

// dynamic shared memory: A (nks+1) x 4 KB | B (nks+1) x 8 KB | sums 256 x ds floats | counts 256 ints
__global__ void __launch_bounds__(KM_ROWS, 2) kmeans_assign_tc_kernel(const KmTcArgs a) {
    extern __shared__ __align__(1024) unsigned char km_smem[];
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ uint32_t s_tmem;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int ds = a.ds, nks = ds >> 3, nks1 = nks + 1, D = a.D, m = blockIdx.y;
    unsigned char *sA = km_smem;
    unsigned char *sB = sA + nks1 * (KM_ROWS * 32);
    float *s_sum = reinterpret_cast<float *>(sB + nks1 * (256 * 32));
    int *s_cnt = reinterpret_cast<int *>(s_sum + 256 * ds);
    const uint32_t a_base = smem_u32(sA), b_base = smem_u32(sB);

    if (tid == 0) { mbar_init(&s_bar, 1); fence_mbar_init(); }
    if (wid == 0) tmem_alloc(&s_tmem, 256);
    // B operand: the subspace's 256 centroids, two rows per thread
    for (int r = tid; r < 256; r += KM_ROWS) {
        const float *src = a.codebook + ((size_t)m * 256 + r) * ds;
        float cn = 0.0f;
        for (int kc = 0; kc < (ds >> 2); ++kc) {
            const float4 v = ldg_f4(src + kc * 4);
            cn = __fmaf_rn(v.x, v.x, cn); cn = __fmaf_rn(v.y, v.y, cn); cn = __fmaf_rn(v.z, v.z, cn); cn = __fmaf_rn(v.w, v.w, cn);
            *reinterpret_cast<float4 *>(sB + (kc >> 1) * (256 * 32) + core_off(r, kc & 1)) = tf32_rna4(v);
        }
        const float ch = tf32_hi(cn);
        unsigned char *x = sB + nks * (256 * 32);
        *reinterpret_cast<float4 *>(x + core_off(r, 0)) = make_float4(ch, cn - ch, 0.0f, 0.0f);
        *reinterpret_cast<float4 *>(x + core_off(r, 1)) = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    }
    for (int i = tid; i < 256 * ds; i += KM_ROWS) s_sum[i] = 0.0f;
    for (int i = tid; i < 256; i += KM_ROWS) s_cnt[i] = 0;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem;
    const uint32_t tlane = tmem + ((uint32_t)(wid * 32) << 16);
    uint32_t phase = 0;
    float sse_local = 0.0f;

    const long long tile0 = (long long)blockIdx.x * KM_TILES;
    for (int t = 0; t < KM_TILES; ++t) {
        const long long r0 = (tile0 + t) * KM_ROWS;
        if (r0 >= a.N) break;                                   // uniform
        const long long i = r0 + tid;
        const bool live = i < a.N;
        const float *x = a.X + (size_t)((live ? i : r0) * a.stride) * D + (size_t)m * ds;
        float xn = 0.0f;
        for (int kc = 0; kc < (ds >> 2); ++kc) {
            float4 v = ldg_f4(x + kc * 4);
            xn = __fmaf_rn(v.x, v.x, xn); xn = __fmaf_rn(v.y, v.y, xn); xn = __fmaf_rn(v.z, v.z, xn); xn = __fmaf_rn(v.w, v.w, xn);
            v.x *= -2.0f; v.y *= -2.0f; v.z *= -2.0f; v.w *= -2.0f;
            *reinterpret_cast<float4 *>(sA + (kc >> 1) * (KM_ROWS * 32) + core_off(tid, kc & 1)) = tf32_rna4(v);
        }
        {
            unsigned char *e = sA + nks * (KM_ROWS * 32);
            *reinterpret_cast<float4 *>(e + core_off(tid, 0)) = make_float4(1.0f, 1.0f, 0.0f, 0.0f);
            *reinterpret_cast<float4 *>(e + core_off(tid, 1)) = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        }
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
        float best = __int_as_float(0x7f800000);
        int bi = 0;
#pragma unroll 4
        for (int ch = 0; ch < 16; ++ch) {
            float v[16];
            tmem_ld16(tlane + (uint32_t)(ch * 16), v);
            tmem_ld_wait();
#pragma unroll
            for (int q = 0; q < 16; ++q)
                if (v[q] < best) { best = v[q]; bi = ch * 16 + q; }    // strict <: the lowest index wins a tie, like the exact kernel
        }
        if (live) {
            for (int kc = 0; kc < (ds >> 2); ++kc) {
                const float4 v = ldg_f4(x + kc * 4);
                atomicAdd(&s_sum[bi * ds + kc * 4 + 0], v.x); atomicAdd(&s_sum[bi * ds + kc * 4 + 1], v.y);
                atomicAdd(&s_sum[bi * ds + kc * 4 + 2], v.z); atomicAdd(&s_sum[bi * ds + kc * 4 + 3], v.w);
            }
            atomicAdd(&s_cnt[bi], 1);
            sse_local += fmaxf(best + xn, 0.0f);
        }
    }
    tc_fence_before();
    __syncthreads();
    for (int i = tid; i < 256 * ds; i += KM_ROWS)
        if (s_sum[i] != 0.0f) atomicAdd(&a.sums[(size_t)m * 256 * ds + i], s_sum[i]);
    for (int i = tid; i < 256; i += KM_ROWS)
        if (s_cnt[i]) atomicAdd(&a.counts[m * 256 + i], s_cnt[i]);
    if (a.sse) {
        const float v = warp_sum_butterfly(sse_local);
        if (lane == 0) atomicAdd(a.sse, (double)v);
    }
    if (wid == 0) tmem_dealloc(tmem, 256);
}

}  // namespace

// one Lloyd assignment pass over rows i * stride, i < N (accumulates sums / counts / sse like pq.cu:assign_kernel<true>)
int launch_kmeans_assign_tc(const float *d_X, long long N, int D, int M, long long stride, const float *d_codebook, float *d_sums,
                            int *d_counts, double *d_sse, cudaStream_t s) {
    const int ds = D / M, nks1 = ds / 8 + 1;
    DR_CHECK((ds & 7) == 0, "k-means on tensor cores needs a sub-dimension that is a multiple of 8 (got %d)", ds);
    const int smem = nks1 * (KM_ROWS * 32) + nks1 * (256 * 32) + 256 * ds * 4 + 256 * 4;
    DR_CHECK(smem <= 100 * 1024, "k-means on tensor cores: sub-dimension %d too large", ds);
    DR_CUDA(cudaFuncSetAttribute(kmeans_assign_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    KmTcArgs a;
    a.X = d_X; a.N = N; a.D = D; a.M = M; a.ds = ds; a.stride = stride; a.codebook = d_codebook; a.sums = d_sums; a.counts = d_counts;
    a.sse = d_sse;
    const long long tiles = (N + KM_ROWS - 1) / KM_ROWS;
    dim3 grid((unsigned)((tiles + KM_TILES - 1) / KM_TILES), (unsigned)M);
    kmeans_assign_tc_kernel<<<grid, KM_ROWS, smem, s>>>(a);
    DR_LAUNCHED();
    return 0;
}
