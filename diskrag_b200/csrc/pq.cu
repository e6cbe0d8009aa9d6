// pq.cu — product-quantiser kernels: K2 ADC table build, K3 k-means training / encode / decode, ADC sums.
//
// K2 replaces DiskANNPQ.compute_distance_table (fast_pq.py:294-318).  The table is produced on CUDA
// cores in exactly numpy's fp32 operation order (diff, square, pairwise row reduction) so that the
// traversal that consumes it is bit-identical to the reference's; the job is bound by writing
// B x M x 1 KiB of table to HBM, not by its 3*256*D flops per query, so there is nothing for the
// tensor pipe to win here.
// K3 replaces DiskANNPQ.fit / encode / decode (fast_pq.py:197-292) (sklearn KMeans per subspace).
#include "common.cuh"

// ---------------------------------------------------------------------------------------------------
// numpy float32 pairwise row sum of t_j = (cen[j] - q[j])^2   (oracle.c: np_pairwise_sum_f32)
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ float sqdiff(const float *cen, const float *q, int j) {
    float d = __fsub_rn(cen[j], q[j]);
    return __fmul_rn(d, d);
}

__device__ float np_pairwise_sqdiff(const float *cen, const float *q, int n) {
    if (n < 8) {
        float res = 0.0f;
        for (int i = 0; i < n; ++i) res = __fadd_rn(res, sqdiff(cen, q, i));
        return res;
    } else if (n <= 128) {
        float r[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) r[j] = sqdiff(cen, q, j);
        int i;
        for (i = 8; i < n - (n % 8); i += 8) {
#pragma unroll
            for (int j = 0; j < 8; ++j) r[j] = __fadd_rn(r[j], sqdiff(cen, q, i + j));
        }
        float res = __fadd_rn(__fadd_rn(__fadd_rn(r[0], r[1]), __fadd_rn(r[2], r[3])),
                              __fadd_rn(__fadd_rn(r[4], r[5]), __fadd_rn(r[6], r[7])));
        for (; i < n; ++i) res = __fadd_rn(res, sqdiff(cen, q, i));
        return res;
    } else {
        int n2 = n / 2;
        n2 -= n2 % 8;
        return __fadd_rn(np_pairwise_sqdiff(cen, q, n2), np_pairwise_sqdiff(cen + n2, q + n2, n - n2));
    }
}

// compile-time sub-dimension: centroid in registers, fully unrolled numpy order
template <int DS>
__device__ __forceinline__ float np_pairwise_ct(const float (&cr)[DS], const float *__restrict__ q) {
    float t[DS];
#pragma unroll
    for (int j = 0; j < DS; ++j) {
        float d = __fsub_rn(cr[j], q[j]);
        t[j] = __fmul_rn(d, d);
    }
    if (DS < 8) {
        float res = 0.0f;
#pragma unroll
        for (int i = 0; i < DS; ++i) res = __fadd_rn(res, t[i]);
        return res;
    }
    float r[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = t[j];
#pragma unroll
    for (int i = 8; i < DS - (DS % 8); i += 8) {
#pragma unroll
        for (int j = 0; j < 8; ++j) r[j] = __fadd_rn(r[j], t[i + j]);
    }
    float res = __fadd_rn(__fadd_rn(__fadd_rn(r[0], r[1]), __fadd_rn(r[2], r[3])),
                          __fadd_rn(__fadd_rn(r[4], r[5]), __fadd_rn(r[6], r[7])));
#pragma unroll
    for (int i = DS - (DS % 8); i < DS; ++i) res = __fadd_rn(res, t[i]);
    return res;
}

// grid (M, ceil(B/QT)); 256 threads: thread c owns centroid c of subspace m and walks the query tile.
// MODE 0: write the fp32 table  out[b][m][c]                                   (reference format)
// MODE 1: per (b, m) min over c -> mn[b][m]; per b max range -> range_bits[b]  (first pass of the u8 table)
// MODE 2: q = min(255, rint((T - mn[b][m]) / scale[b])) -> out8[b][m][c]       (second pass of the u8 table)
#define LUT_QT 64
template <int DS, int MODE>  // DS == 0: runtime sub-dimension
__global__ void __launch_bounds__(256) lut_kernel(const float *__restrict__ codebook, const float *__restrict__ Q,
                                                  long long B, int D, int M, float *__restrict__ out,
                                                  float *__restrict__ mn, unsigned *__restrict__ range_bits,
                                                  const float *__restrict__ scale, uint8_t *__restrict__ out8) {
    extern __shared__ float s_qs[];  // [LUT_QT][ds]
    __shared__ float s_lo[8][LUT_QT], s_hi[8][LUT_QT];
    const int m = blockIdx.x, ds = DS ? DS : D / M, c = threadIdx.x;
    const long long b0 = (long long)blockIdx.y * LUT_QT;
    const int nb = (int)((B - b0 < LUT_QT) ? (B - b0) : LUT_QT);
    for (int i = threadIdx.x; i < nb * ds; i += blockDim.x) {
        int bb = i / ds, j = i - bb * ds;
        s_qs[i] = __ldg(Q + (size_t)(b0 + bb) * D + m * ds + j);
    }
    __syncthreads();
    const float *cen = codebook + ((size_t)m * 256 + c) * ds;
    float cr[DS ? DS : 1];
    if (DS) {
#pragma unroll
        for (int j = 0; j < (DS ? DS : 1); ++j) cr[j] = __ldg(cen + j);
    }
    for (int bb = 0; bb < nb; ++bb) {
        float v;
        if (DS) v = np_pairwise_ct<(DS ? DS : 1)>(cr, s_qs + bb * (DS ? DS : 1));
        else v = np_pairwise_sqdiff(cen, s_qs + bb * ds, ds);
        if (MODE == 0) {
            out[((size_t)(b0 + bb) * M + m) * 256 + c] = v;
        } else if (MODE == 1) {
            float lo = v, hi = v;
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1) {
                lo = fminf(lo, __shfl_xor_sync(DR_FULL, lo, off));
                hi = fmaxf(hi, __shfl_xor_sync(DR_FULL, hi, off));
            }
            if ((c & 31) == 0) { s_lo[c >> 5][bb] = lo; s_hi[c >> 5][bb] = hi; }
        } else {
            const float lo = mn[(size_t)(b0 + bb) * M + m];
            float qv = rintf(__fdiv_rn(__fsub_rn(v, lo), scale[b0 + bb]));
            out8[((size_t)(b0 + bb) * M + m) * 256 + c] = (uint8_t)(qv > 255.0f ? 255.0f : qv);
        }
    }
    if (MODE == 1) {
        __syncthreads();
        if (c < nb) {
            float lo = s_lo[0][c], hi = s_hi[0][c];
#pragma unroll
            for (int w = 1; w < 8; ++w) { lo = fminf(lo, s_lo[w][c]); hi = fmaxf(hi, s_hi[w][c]); }
            mn[(size_t)(b0 + c) * M + m] = lo;
            atomicMax(range_bits + b0 + c, __float_as_uint(__fsub_rn(hi, lo)));  // non-negative floats order like uints
        }
    }
}

template <int MODE>
static int lut_launch(const float *d_codebook, const float *d_Q, int64_t B, int D, int M, float *d_out, float *d_mn,
                      unsigned *d_range, const float *d_scale, uint8_t *d_out8, cudaStream_t s) {
    const int ds = D / M;
    const size_t smem = (size_t)LUT_QT * ds * sizeof(float);
    DR_CHECK(smem <= 40 * 1024, "dr_lut_build: sub-dimension %d too large", ds);
    const long long tiles = (B + LUT_QT - 1) / LUT_QT;
    for (long long t0 = 0; t0 < tiles; t0 += 65535) {  // gridDim.y limit
        long long nt = tiles - t0 < 65535 ? tiles - t0 : 65535;
        dim3 grid(M, (unsigned)nt);
        const size_t qo = (size_t)t0 * LUT_QT;
        const float *q = d_Q + qo * D;
        float *o = d_out ? d_out + qo * M * 256 : nullptr;
        float *mnp = d_mn ? d_mn + qo * M : nullptr;
        unsigned *rg = d_range ? d_range + qo : nullptr;
        const float *sc = d_scale ? d_scale + qo : nullptr;
        uint8_t *o8 = d_out8 ? d_out8 + qo * M * 256 : nullptr;
        long long bb = B - t0 * LUT_QT;
        switch (ds) {
#define LUT_CASE(X) case X: lut_kernel<X, MODE><<<grid, 256, smem, s>>>(d_codebook, q, bb, D, M, o, mnp, rg, sc, o8); break;
            LUT_CASE(4) LUT_CASE(8) LUT_CASE(12) LUT_CASE(16) LUT_CASE(24) LUT_CASE(32) LUT_CASE(48) LUT_CASE(64)
#undef LUT_CASE
            default: lut_kernel<0, MODE><<<grid, 256, smem, s>>>(d_codebook, q, bb, D, M, o, mnp, rg, sc, o8); break;
        }
        DR_LAUNCHED();
    }
    return 0;
}

int launch_lut_build(const float *d_codebook, const float *d_Q, int64_t B, int D, int M, float *d_out, cudaStream_t s) {
    DR_CHECK(M > 0 && D % M == 0, "dr_lut_build: D=%d not divisible by M=%d", D, M);
    if (B == 0) return 0;
    return lut_launch<0>(d_codebook, d_Q, B, D, M, d_out, nullptr, nullptr, nullptr, nullptr, s);
}

// ---------------------------------------------------------------------------------------------------
// u8 table for the throughput mode.  The table entry only has to rank centroids inside one subspace, so it is
// built from  t[c] = ||c||^2 - 2 q_m . c  (the ||q_m||^2 term is constant per subspace and is folded into the
// per-query offset): one fmaf chain per entry instead of numpy's sub/mul/pairwise-add order.
//   pass 1 (thread = query): lo[b][m] = min_c t, range[b] = max_m (max_c t - lo)
//   finalize               : scale[b] = range / 255 (1 if 0), offset[b] = sum_m lo[b][m] + ||q||^2
//   pass 2 (thread = centroid): out8[b][m][c] = clamp(rint(fma(t, inv, c0)), 0, 255), inv = 1/scale[b], c0 = -(lo[b][m] * inv)
// Restated bit-for-bit by oracle.c:orc_lut_u8.
// ---------------------------------------------------------------------------------------------------
#define U8_QT 256
template <int DS>
__global__ void __launch_bounds__(256) lut_u8_stats_kernel(const float *__restrict__ codebook, const float *__restrict__ Q,
                                                           long long B, int D, int M, float *__restrict__ lo_out,
                                                           unsigned *__restrict__ range_bits) {
    __shared__ __align__(16) float s_c[256 * DS];
    __shared__ float s_cn[256];
    const int m = blockIdx.x;
    const long long b = (long long)blockIdx.y * U8_QT + threadIdx.x;
    const float *cb = codebook + (size_t)m * 256 * DS;
    for (int i = threadIdx.x; i < 256 * DS; i += 256) s_c[i] = __ldg(cb + i);
    __syncthreads();
    {
        float acc = 0.0f;
#pragma unroll
        for (int j = 0; j < DS; ++j) acc = __fmaf_rn(s_c[threadIdx.x * DS + j], s_c[threadIdx.x * DS + j], acc);
        s_cn[threadIdx.x] = acc;
    }
    __syncthreads();
    if (b >= B) return;
    float q2[DS];
#pragma unroll
    for (int j = 0; j < DS; ++j) q2[j] = -2.0f * __ldg(Q + (size_t)b * D + m * DS + j);
    float lo = __int_as_float(0x7f800000), hi = -__int_as_float(0x7f800000);
#pragma unroll 4
    for (int c = 0; c < 256; ++c) {
        float acc = s_cn[c];
#pragma unroll
        for (int j = 0; j < DS; ++j) acc = __fmaf_rn(q2[j], s_c[c * DS + j], acc);
        lo = fminf(lo, acc);
        hi = fmaxf(hi, acc);
    }
    lo_out[(size_t)b * M + m] = lo;
    atomicMax(range_bits + b, __float_as_uint(__fsub_rn(hi, lo)));  // a non-negative float orders like its bits
}

__global__ void lut_u8_finalize_kernel(const float *__restrict__ lo, const unsigned *__restrict__ range_bits,
                                       const float *__restrict__ Q, long long B, int D, int M,
                                       float *__restrict__ scale, float *__restrict__ offset) {
    long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    float range = __uint_as_float(range_bits[b]);
    scale[b] = range > 0.0f ? __fdiv_rn(range, 255.0f) : 1.0f;
    float acc = 0.0f;
    for (int m = 0; m < M; ++m) acc = __fadd_rn(acc, lo[(size_t)b * M + m]);
    float qn = 0.0f;
    for (int j = 0; j < D; ++j) { float v = Q[(size_t)b * D + j]; qn = __fmaf_rn(v, v, qn); }
    offset[b] = __fadd_rn(acc, qn);
}

#define U8_QT2 64
// thread = centroid c of subspace m, walking a 64-query tile (coalesced 256 B rows)
template <int DS>
__global__ void __launch_bounds__(256) lut_u8_quant_kernel(const float *__restrict__ codebook, const float *__restrict__ Q,
                                                           long long B, int D, int M, const float *__restrict__ lo_in,
                                                           const float *__restrict__ scale, uint8_t *__restrict__ out8) {
    __shared__ __align__(16) float s_q2[U8_QT2 * DS];
    __shared__ float s_lo[U8_QT2], s_sc[U8_QT2];
    const int m = blockIdx.x, c = threadIdx.x;
    const long long b0 = (long long)blockIdx.y * U8_QT2;
    const int nb = (int)((B - b0 < U8_QT2) ? (B - b0) : U8_QT2);
    for (int i = threadIdx.x; i < nb * DS; i += 256) {
        int bb = i / DS, j = i - bb * DS;
        s_q2[i] = -2.0f * __ldg(Q + (size_t)(b0 + bb) * D + m * DS + j);
    }
    if (threadIdx.x < nb) {   // q = rint(t * inv + c0), inv = 1 / scale, c0 = -(lo * inv)
        const float inv = __fdiv_rn(1.0f, scale[b0 + threadIdx.x]);
        s_sc[threadIdx.x] = inv;
        s_lo[threadIdx.x] = -__fmul_rn(lo_in[(size_t)(b0 + threadIdx.x) * M + m], inv);
    }
    float cr[DS];
    float cn = 0.0f;
#pragma unroll
    for (int j = 0; j < DS; ++j) {
        cr[j] = __ldg(codebook + ((size_t)m * 256 + c) * DS + j);
        cn = __fmaf_rn(cr[j], cr[j], cn);
    }
    __syncthreads();
    uint8_t *o = out8 + ((size_t)b0 * M + m) * 256 + c;
    for (int bb = 0; bb < nb; ++bb) {
        float acc = cn;
#pragma unroll
        for (int j = 0; j < DS; ++j) acc = __fmaf_rn(s_q2[bb * DS + j], cr[j], acc);
        uint32_t qv;
        asm("cvt.rni.sat.u8.f32 %0, %1;" : "=r"(qv) : "f"(__fmaf_rn(acc, s_sc[bb], s_lo[bb])));   // round-half-even, clamp to [0, 255]
        o[(size_t)bb * M * 256] = (uint8_t)qv;
    }
}

int launch_lut_u8_finalize(const float *d_lo, const unsigned *d_range, const float *d_Q, int64_t B, int D, int M, float *d_scale,
                           float *d_offset, cudaStream_t s) {
    if (B == 0) return 0;
    lut_u8_finalize_kernel<<<(unsigned)((B + 127) / 128), 128, 0, s>>>(d_lo, d_range, d_Q, B, D, M, d_scale, d_offset);
    DR_LAUNCHED();
    return 0;
}

// Word layout for the bank-per-lane search table (search_fast.cu): out8[b][w][c][j] = entry of subspace 4w + j,
// centroid c.  Block = (code word w, 64-query tile); thread c computes the four subspaces' entries of its centroid
// and stores one packed 32-bit word per query: a warp writes 128 contiguous bytes.  Same arithmetic, entry for
// entry, as lut_u8_quant_kernel.
template <int DS>
__global__ void __launch_bounds__(256) lut_u8_quant_word_kernel(const float *__restrict__ codebook, const float *__restrict__ Q,
                                                                long long B, int D, int M, const float *__restrict__ lo_in,
                                                                const float *__restrict__ scale, uint32_t *__restrict__ out32) {
    __shared__ __align__(16) float s_q2[U8_QT2 * 4 * DS];
    __shared__ float s_lo[U8_QT2 * 4], s_sc[U8_QT2];
    const int w = blockIdx.x, c = threadIdx.x, words = M >> 2;
    const long long b0 = (long long)blockIdx.y * U8_QT2;
    const int nb = (int)((B - b0 < U8_QT2) ? (B - b0) : U8_QT2);
    for (int i = threadIdx.x; i < nb * 4 * DS; i += 256) {
        int bb = i / (4 * DS), j = i - bb * (4 * DS);
        s_q2[i] = -2.0f * __ldg(Q + (size_t)(b0 + bb) * D + (size_t)w * 4 * DS + j);
    }
    if (threadIdx.x < nb) s_sc[threadIdx.x] = __fdiv_rn(1.0f, scale[b0 + threadIdx.x]);
    __syncthreads();
    if (threadIdx.x < nb * 4) {   // c0 = -(lo * inv) per (query, subspace of this word)
        const int bb = threadIdx.x >> 2, j = threadIdx.x & 3;
        s_lo[threadIdx.x] = -__fmul_rn(lo_in[(size_t)(b0 + bb) * M + 4 * w + j], s_sc[bb]);
    }
    float cr[4][DS], cn[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        cn[j] = 0.0f;
#pragma unroll
        for (int i = 0; i < DS; ++i) {
            cr[j][i] = __ldg(codebook + ((size_t)(4 * w + j) * 256 + c) * DS + i);
            cn[j] = __fmaf_rn(cr[j][i], cr[j][i], cn[j]);
        }
    }
    __syncthreads();
    uint32_t *o = out32 + ((size_t)b0 * words + w) * 256 + c;
    for (int bb = 0; bb < nb; ++bb) {
        uint32_t packed = 0u;
        const float inv = s_sc[bb];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float acc = cn[j];
#pragma unroll
            for (int i = 0; i < DS; ++i) acc = __fmaf_rn(s_q2[(bb * 4 + j) * DS + i], cr[j][i], acc);
            uint32_t qv;
            asm("cvt.rni.sat.u8.f32 %0, %1;" : "=r"(qv) : "f"(__fmaf_rn(acc, inv, s_lo[bb * 4 + j])));
            packed |= qv << (8 * j);
        }
        o[(size_t)bb * words * 256] = packed;
    }
}

template <int DS>
static int lut_u8_launch(const float *d_codebook, const float *d_Q, int64_t B, int D, int M, uint8_t *d_out8, float *d_scale,
                         float *d_offset, float *d_lo, unsigned *d_range, int word_layout, cudaStream_t s) {
    DR_CUDA(cudaMemsetAsync(d_range, 0, (size_t)B * 4, s));
    for (long long t0 = 0; t0 * U8_QT < B; t0 += 65535) {
        long long tiles = (B - t0 * U8_QT + U8_QT - 1) / U8_QT;
        if (tiles > 65535) tiles = 65535;
        dim3 grid(M, (unsigned)tiles);
        const size_t qo = (size_t)t0 * U8_QT;
        lut_u8_stats_kernel<DS><<<grid, 256, 0, s>>>(d_codebook, d_Q + qo * D, B - (long long)qo, D, M, d_lo + qo * M, d_range + qo);
        DR_LAUNCHED();
    }
    lut_u8_finalize_kernel<<<(unsigned)((B + 127) / 128), 128, 0, s>>>(d_lo, d_range, d_Q, B, D, M, d_scale, d_offset);
    DR_LAUNCHED();
    for (long long t0 = 0; t0 * U8_QT2 < B; t0 += 65535) {
        long long tiles = (B - t0 * U8_QT2 + U8_QT2 - 1) / U8_QT2;
        if (tiles > 65535) tiles = 65535;
        const size_t qo = (size_t)t0 * U8_QT2;
        if (word_layout) {
            dim3 grid(M >> 2, (unsigned)tiles);
            lut_u8_quant_word_kernel<DS><<<grid, 256, 0, s>>>(d_codebook, d_Q + qo * D, B - (long long)qo, D, M, d_lo + qo * M,
                                                              d_scale + qo, reinterpret_cast<uint32_t *>(d_out8 + qo * M * 256));
        } else {
            dim3 grid(M, (unsigned)tiles);
            lut_u8_quant_kernel<DS><<<grid, 256, 0, s>>>(d_codebook, d_Q + qo * D, B - (long long)qo, D, M, d_lo + qo * M, d_scale + qo,
                                                         d_out8 + qo * M * 256);
        }
        DR_LAUNCHED();
    }
    return 0;
}

// d_out8: word_layout == 0: u8[B][M][256];  word_layout == 1 (M % 4 == 0): u8[B][M/4][256][4] (see above);
// d_scale/d_offset f32[B]; d_mn f32[B][M] and d_range u32[B] are scratch
int launch_lut_build_u8(const float *d_codebook, const float *d_Q, int64_t B, int D, int M, uint8_t *d_out8, float *d_scale,
                        float *d_offset, float *d_mn, unsigned *d_range, int word_layout, cudaStream_t s) {
    DR_CHECK(M > 0 && D % M == 0, "dr_lut_build: D=%d not divisible by M=%d", D, M);
    DR_CHECK(!word_layout || (M & 3) == 0, "dr_lut_build(u8): the word layout needs M %% 4 == 0 (M=%d)", M);
    if (B == 0) return 0;
    switch (D / M) {
#define U8_CASE(X) case X: return lut_u8_launch<X>(d_codebook, d_Q, B, D, M, d_out8, d_scale, d_offset, d_mn, d_range, word_layout, s);
        U8_CASE(1) U8_CASE(2) U8_CASE(3) U8_CASE(4) U8_CASE(5) U8_CASE(6) U8_CASE(8) U8_CASE(12) U8_CASE(16) U8_CASE(24) U8_CASE(32)
#undef U8_CASE
        default: break;
    }
    dr_set_error("dr_search(u8): sub-dimension %d is not instantiated (1-6, 8, 12, 16, 24, 32)", D / M);
    return 2;
}

// word layout [b][w][c][4] -> plain [b][m][c] (tests only)
__global__ void lut_u8_unpermute_kernel(const uint8_t *__restrict__ in, long long total, int M, uint8_t *__restrict__ out) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i & 255);
        const long long bm = i >> 8;
        const int m = (int)(bm % M);
        const long long b = bm / M;
        out[i] = in[((b * (M >> 2) + (m >> 2)) * 256 + c) * 4 + (m & 3)];
    }
}
int launch_lut_u8_unpermute(const uint8_t *d_words, int64_t B, int M, uint8_t *d_plain, cudaStream_t s) {
    if (B == 0) return 0;
    const long long total = (long long)B * M * 256;
    lut_u8_unpermute_kernel<<<(unsigned)((total + 255) / 256 < 65535 ? (total + 255) / 256 : 65535), 256, 0, s>>>(d_words, total, M, d_plain);
    DR_LAUNCHED();
    return 0;
}

// ---------------------------------------------------------------------------------------------------
// ADC sums, sequential order (fast_pq.py:320-328): one thread per code row
// ---------------------------------------------------------------------------------------------------
__global__ void adc_kernel(const uint8_t *__restrict__ codes, const float *__restrict__ lut, long long n, int M,
                           float *__restrict__ out) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint8_t *code = codes + (size_t)i * M;
    float acc = 0.0f;
    for (int m = 0; m < M; ++m) acc = __fadd_rn(acc, __ldg(lut + m * 256 + code[m]));
    out[i] = acc;
}

int launch_adc(const uint8_t *d_codes, const float *d_lut, int64_t n, int M, float *d_out, cudaStream_t s) {
    if (n == 0) return 0;
    adc_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(d_codes, d_lut, n, M, d_out);
    DR_LAUNCHED();
    return 0;
}

// ---------------------------------------------------------------------------------------------------
// K3: nearest-centroid assignment (encode) and Lloyd k-means
//   grid (ceil(N/256), M), 256 threads: the CTA stages subspace m's 256 centroids in shared memory,
//   each thread owns one point's sub-vector in registers and scans the centroids (broadcast reads).
//   Lowest index wins ties (strict <), as sklearn's argmin does.
// ---------------------------------------------------------------------------------------------------
#define PQ_MAX_DS 64
template <bool ACCUM>
__global__ void __launch_bounds__(256) assign_kernel(const float *__restrict__ X, long long N, int D, int M,
                                                     long long stride, long long offset,
                                                     const float *__restrict__ codebook, uint8_t *__restrict__ codes,
                                                     float *__restrict__ sums, int *__restrict__ counts,
                                                     double *__restrict__ sse) {
    extern __shared__ float s_pq[];  // centroids [256][ds]; ACCUM: + sums [256][ds] + counts [256]
    const int m = blockIdx.y, ds = D / M;
    float *s_cen = s_pq;
    float *s_sum = s_pq + 256 * ds;
    int *s_cnt = reinterpret_cast<int *>(s_sum + 256 * ds);
    for (int i = threadIdx.x; i < 256 * ds; i += blockDim.x) {
        s_cen[i] = __ldg(codebook + (size_t)m * 256 * ds + i);
        if (ACCUM) s_sum[i] = 0.0f;
    }
    if (ACCUM) for (int i = threadIdx.x; i < 256; i += blockDim.x) s_cnt[i] = 0;
    __syncthreads();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    float best = 0.0f;
    if (i < N) {
        const long long row = i * stride + offset;  // training subsample walks every stride-th row
        const float *x = X + (size_t)row * D + m * ds;
        float xr[PQ_MAX_DS];
        for (int j = 0; j < ds; ++j) xr[j] = __ldg(x + j);
        int bi = 0;
        best = __int_as_float(0x7f800000);
        for (int c = 0; c < 256; ++c) {
            const float *cen = s_cen + c * ds;
            float acc = 0.0f;
            for (int j = 0; j < ds; ++j) {
                float d = __fsub_rn(xr[j], cen[j]);
                acc = __fmaf_rn(d, d, acc);
            }
            if (acc < best) { best = acc; bi = c; }
        }
        if (codes) codes[(size_t)i * M + m] = (uint8_t)bi;
        if (ACCUM) {
            for (int j = 0; j < ds; ++j) atomicAdd(&s_sum[bi * ds + j], xr[j]);
            atomicAdd(&s_cnt[bi], 1);
        }
    }
    if (ACCUM) {
        __syncthreads();
        for (int t = threadIdx.x; t < 256 * ds; t += blockDim.x)
            if (s_sum[t] != 0.0f) atomicAdd(&sums[(size_t)m * 256 * ds + t], s_sum[t]);
        for (int t = threadIdx.x; t < 256; t += blockDim.x)
            if (s_cnt[t]) atomicAdd(&counts[m * 256 + t], s_cnt[t]);
        if (sse) {
            float v = (i < N) ? best : 0.0f;
            v = warp_sum_butterfly(v);
            if ((threadIdx.x & 31) == 0) atomicAdd(sse, (double)v);
        }
    }
}

// new centroid = mean of its members; an empty cluster is re-seeded from a pseudo-random training row
__global__ void update_kernel(float *__restrict__ codebook, const float *__restrict__ sums, const int *__restrict__ counts,
                              const float *__restrict__ X, long long N, int D, int M, unsigned long long seed, int iter) {
    const int ds = D / M;
    const int mc = blockIdx.x;  // m*256 + c
    const int cnt = counts[mc];
    const int m = mc >> 8;
    for (int j = threadIdx.x; j < ds; j += blockDim.x) {
        float v;
        if (cnt > 0) v = sums[(size_t)mc * ds + j] / (float)cnt;
        else {
            unsigned long long z = seed + 0x9E3779B97F4A7C15ull * (unsigned long long)(mc * 131 + iter * 7919 + 1);
            z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; z ^= z >> 31;
            v = X[(size_t)(z % (unsigned long long)N) * D + m * ds + j];
        }
        codebook[(size_t)mc * ds + j] = v;
    }
}

__global__ void init_codebook_kernel(float *__restrict__ codebook, const float *__restrict__ X, long long N, int D, int M,
                                     unsigned long long seed) {
    const int ds = D / M;
    const int mc = blockIdx.x, m = mc >> 8, c = mc & 255;
    // 256 distinct rows per subspace: an affine walk with an odd stride over N (distinct while 256*step < N wraps rarely)
    unsigned long long z = seed + 0x9E3779B97F4A7C15ull * (unsigned long long)(m + 1);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; z ^= z >> 31;
    unsigned long long step = (unsigned long long)N / 256ull;
    if (step == 0) step = 1;
    unsigned long long row = (z % (unsigned long long)N + (unsigned long long)c * step) % (unsigned long long)N;
    for (int j = threadIdx.x; j < ds; j += blockDim.x) codebook[(size_t)mc * ds + j] = X[(size_t)row * D + m * ds + j];
}

// ---- k-means++ seeding (sklearn's KMeans(init="k-means++") at fast_pq.py:232-240; Arthur & Vassilvitskii 2007) -------------
// All M subspaces at once, one centroid per round: every (row, subspace) keeps its squared distance to the nearest centre chosen
// so far; the next centre of a subspace is a row drawn with probability proportional to that distance, drawn as the argmax of
// mind_i / E_i with E_i ~ Exp(1) (the exponential-race form of weighted sampling: one max-reduction, no prefix sums).
__device__ __forceinline__ float kpp_exp1(unsigned long long seed, int round, int m, long long i) {
    unsigned long long z = seed + 0x9E3779B97F4A7C15ull * ((unsigned long long)i * 257ull + (unsigned long long)m * 65537ull + (unsigned long long)round + 1ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; z ^= z >> 31;
    const float u = ((float)(unsigned)(z >> 40) + 0.5f) * (1.0f / 16777216.0f);     // (0, 1)
    return -__logf(u);
}
// grid (ceil(n / 256), M): distances of every row to the centre chosen in the previous round, running minimum, race key
__global__ void __launch_bounds__(256) kpp_round_kernel(const float *__restrict__ X, long long n, long long stride, int D, int M,
                                                        const float *__restrict__ codebook, int round, float *__restrict__ mind,
                                                        unsigned long long *__restrict__ best, unsigned long long seed) {
    const int ds = D / M, m = blockIdx.y;
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    unsigned long long key = 0ull;
    if (i < n) {
        const float *x = X + (size_t)(i * stride) * D + (size_t)m * ds;
        const float *c = codebook + ((size_t)m * 256 + (round - 1)) * ds;         // the centre picked last
        float d = 0.0f;
        for (int j = 0; j < ds; ++j) { const float t = x[j] - __ldg(c + j); d = __fmaf_rn(t, t, d); }
        float md = d;
        if (round > 1) md = fminf(md, mind[(size_t)m * n + i]);
        mind[(size_t)m * n + i] = md;
        const float r = md / kpp_exp1(seed, round, m, i);
        key = ((unsigned long long)__float_as_uint(r) << 32) | (unsigned long long)(unsigned)i;   // r >= 0: bits order like the value
    }
    // block max, one atomic per block
    for (int off = 16; off >= 1; off >>= 1) { const unsigned long long o = __shfl_xor_sync(DR_FULL, key, off); key = o > key ? o : key; }
    __shared__ unsigned long long s_k[8];
    if ((threadIdx.x & 31) == 0) s_k[threadIdx.x >> 5] = key;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) key = s_k[w] > key ? s_k[w] : key;
        atomicMax(best + m, key);
    }
}
// M blocks: centre `round` of subspace m = the winning row (round 0: a uniformly drawn row); resets the race
__global__ void kpp_pick_kernel(const float *__restrict__ X, long long n, long long stride, int D, int M, float *__restrict__ codebook,
                                int round, unsigned long long *__restrict__ best, unsigned long long seed) {
    const int ds = D / M, m = blockIdx.x;
    long long row;
    if (round == 0) {
        unsigned long long z = seed + 0x9E3779B97F4A7C15ull * (unsigned long long)(m + 1);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; z ^= z >> 31;
        row = (long long)(z % (unsigned long long)n);
    } else {
        row = (long long)(best[m] & 0xFFFFFFFFull);
    }
    for (int j = threadIdx.x; j < ds; j += blockDim.x) codebook[((size_t)m * 256 + round) * ds + j] = X[(size_t)(row * stride) * D + m * ds + j];
    __syncthreads();
    if (threadIdx.x == 0) best[m] = 0ull;
}
static std::atomic<int> g_kmeans_pp{1};
void pq_train_set_kmeanspp(int enable) { g_kmeans_pp.store(enable ? 1 : 0); }

static size_t assign_smem(int ds, bool accum) { return (size_t)256 * ds * 4 * (accum ? 2 : 1) + (accum ? 1024 : 0); }

int launch_pq_encode(const float *d_codebook, const float *d_X, int64_t N, int D, int M, uint8_t *d_codes, cudaStream_t s) {
    DR_CHECK(M > 0 && D % M == 0, "dr_pq_encode: D=%d not divisible by M=%d", D, M);
    const int ds = D / M;
    DR_CHECK(ds <= PQ_MAX_DS, "dr_pq_encode: sub-dimension %d > %d not supported", ds, PQ_MAX_DS);
    if (N == 0) return 0;
    size_t smem = assign_smem(ds, false);
    DR_CUDA(cudaFuncSetAttribute(assign_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)((N + 255) / 256), M);
    assign_kernel<false><<<grid, 256, smem, s>>>(d_X, N, D, M, 1, 0, d_codebook, d_codes, nullptr, nullptr, nullptr);
    DR_LAUNCHED();
    return 0;
}

int launch_kmeans_assign_tc(const float *d_X, long long N, int D, int M, long long stride, const float *d_codebook, float *d_sums,
                            int *d_counts, double *d_sse, cudaStream_t s);   // kmeans_tc.cu (tcgen05)
static std::atomic<int> g_kmeans_tc{1};
void pq_train_set_tensor_cores(int enable) { g_kmeans_tc.store(enable ? 1 : 0); }

int launch_pq_train(const float *d_X, int64_t N, int D, int M, int iters, uint64_t seed, float *d_codebook,
                    double *out_mse, cudaStream_t s) {
    DR_CHECK(M > 0 && D % M == 0, "dr_pq_train: D=%d not divisible by M=%d", D, M);
    DR_CHECK(N >= 256, "dr_pq_train: need at least 256 training vectors (got %lld)", (long long)N);
    const int ds = D / M;
    DR_CHECK(ds <= PQ_MAX_DS, "dr_pq_train: sub-dimension %d > %d not supported", ds, PQ_MAX_DS);
    if (iters <= 0) iters = 25;
    // training subsample: every stride-th row (sklearn trains on all rows; 1024 rows per centroid is plenty)
    long long ntrain = N, stride = 1;
    const long long cap = 262144;
    if (N > cap) { stride = N / cap; ntrain = N / stride; }
    float *d_sums = nullptr; int *d_counts = nullptr; double *d_sse = nullptr;
    DR_CUDA(cudaMalloc(&d_sums, (size_t)M * 256 * ds * 4));
    DR_CUDA(cudaMalloc(&d_counts, (size_t)M * 256 * 4));
    DR_CUDA(cudaMalloc(&d_sse, 8));
    if (g_kmeans_pp.load()) {
        // k-means++ over a subsample of at most 65536 training rows (every kstride-th training row)
        long long kn = ntrain, kstride = stride;
        if (kn > 65536) { const long long f = kn / 65536; kstride = stride * f; kn = ntrain / f; }
        float *d_mind = nullptr; unsigned long long *d_best = nullptr;
        DR_CUDA(cudaMalloc(&d_mind, (size_t)M * kn * 4));
        DR_CUDA(cudaMalloc(&d_best, (size_t)M * 8));
        DR_CUDA(cudaMemsetAsync(d_best, 0, (size_t)M * 8, s));
        dim3 kgrid((unsigned)((kn + 255) / 256), M);
        for (int round = 0; round < 256; ++round) {
            if (round > 0) {
                kpp_round_kernel<<<kgrid, 256, 0, s>>>(d_X, kn, kstride, D, M, d_codebook, round, d_mind, d_best, seed);
                DR_LAUNCHED();
            }
            kpp_pick_kernel<<<M, 32, 0, s>>>(d_X, kn, kstride, D, M, d_codebook, round, d_best, seed);
            DR_LAUNCHED();
        }
        DR_CUDA(cudaStreamSynchronize(s));
        cudaFree(d_mind); cudaFree(d_best);
    } else {
        init_codebook_kernel<<<M * 256, 32, 0, s>>>(d_codebook, d_X, N, D, M, seed);
        DR_LAUNCHED();
    }
    size_t smem = assign_smem(ds, true);
    DR_CUDA(cudaFuncSetAttribute(assign_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)((ntrain + 255) / 256), M);
    const bool use_tc = g_kmeans_tc.load() != 0 && (ds & 7) == 0 && ds <= 64;
    for (int it = 0; it < iters; ++it) {
        DR_CUDA(cudaMemsetAsync(d_sums, 0, (size_t)M * 256 * ds * 4, s));
        DR_CUDA(cudaMemsetAsync(d_counts, 0, (size_t)M * 256 * 4, s));
        DR_CUDA(cudaMemsetAsync(d_sse, 0, 8, s));
        // assignment: tensor cores (TF32 contraction rows x centroids) when the sub-dimension allows, else CUDA cores;
        // the last pass, which only measures the error of the final codebook, is always exact
        if (use_tc && it + 1 < iters) {
            if (launch_kmeans_assign_tc(d_X, ntrain, D, M, stride, d_codebook, d_sums, d_counts, d_sse, s)) return 1;
        } else {
            assign_kernel<true><<<grid, 256, smem, s>>>(d_X, ntrain, D, M, stride, 0, d_codebook, nullptr, d_sums, d_counts, d_sse);
            DR_LAUNCHED();
        }
        if (it + 1 < iters) {  // the last pass only measures the error of the final codebook
            update_kernel<<<M * 256, 32, 0, s>>>(d_codebook, d_sums, d_counts, d_X, N, D, M, seed, it);
            DR_LAUNCHED();
        }
    }
    if (out_mse) {
        double sse = 0.0;
        DR_CUDA(cudaMemcpyAsync(&sse, d_sse, 8, cudaMemcpyDeviceToHost, s));
        DR_CUDA(cudaStreamSynchronize(s));
        *out_mse = sse / ((double)ntrain * (double)D);
    } else {
        DR_CUDA(cudaStreamSynchronize(s));
    }
    cudaFree(d_sums); cudaFree(d_counts); cudaFree(d_sse);
    return 0;
}

// decode (fast_pq.py:269-292): gather centroids
__global__ void decode_kernel(const float *__restrict__ codebook, const uint8_t *__restrict__ codes, long long N, int D, int M,
                              float *__restrict__ out) {
    const int ds = D / M;
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= N * D) return;
    long long i = t / D;
    int j = (int)(t - i * D), m = j / ds, jj = j - m * ds;
    out[t] = __ldg(codebook + ((size_t)m * 256 + codes[(size_t)i * M + m]) * ds + jj);
}

int launch_pq_decode(const float *d_codebook, const uint8_t *d_codes, int64_t N, int D, int M, float *d_out, cudaStream_t s) {
    if (N == 0) return 0;
    long long tot = (long long)N * D;
    decode_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, s>>>(d_codebook, d_codes, N, D, M, d_out);
    DR_LAUNCHED();
    return 0;
}
