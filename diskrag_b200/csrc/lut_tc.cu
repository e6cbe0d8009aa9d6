// lut_tc.cu — K2 on the 5th-generation tensor cores: the 8-bit ADC table of the throughput search, built with
// tcgen05.mma (kind::tf32, accumulators in TMEM) instead of CUDA-core FMA chains.
//
// Replaces DiskANNPQ.compute_distance_table (pydiskann/pq/fast_pq.py:294-318) for the u8 table of search_fast.cu:
//   t[b][m][c]   = ||C[m][c]||^2 - 2 q_b,m . C[m][c]                 (the ||q_m||^2 term goes into the per-query offset)
//   lo[b][m]     = min_c t,   scale[b] = max_m (max_c t - lo) / 255,  offset[b] = sum_m lo[b][m] + ||q_b||^2
//   out[b][m][c] = sat_u8(rint((t - lo[b][m]) / scale[b]))
// which is the contract of pq.cu:launch_lut_build_u8 (restated by oracle.c:orc_lut_u8), here with the inner products
// in TF32 (10-bit mantissa inputs, fp32 accumulation): an entry may differ from the exact table by one unit
// (tests/test_lut_tc_gpu.py bounds it); traversal quality is unchanged (recall checked in the same test).
//
// Shape: queries x centroids is a genuine dense contraction, one per subspace: [128 queries x ds] . [ds x 256].
// One CTA = one 128-query tile; a "stage" = one code word (4 subspaces) x one quarter of the centroids (64):
// four MMAs D_j[128 x 64] (+)= A_j[128 x 8] . B_j[64 x 8]^T per K-step write 4 x 64 = 256 TMEM columns, and the 128
// threads (thread = TMEM lane = query) read them back with tcgen05.ld.  The affine part of the map is folded into
// one extra K-step so that the epilogue is a bare convert-and-pack:
//   phase 1 (statistics):  A = [-2 q | 1, 1, 0..],            B = [c | cn_hi, cn_lo, 0..]          -> D = t
//   phase 2 (quantise)  :  A = [-2 inv q | inv_hi, inv_hi, inv_lo, k_hi, k_lo, 0..],
//                          B = [c | cn_hi, cn_lo, cn_hi, 1, 1, 0..]                                -> D = t * inv + k
//   with inv = 1 / scale[b], k = -lo[b][m] * inv, and x_hi / x_lo the exact two-term TF32 split of x.
// Operands are staged in shared memory in the canonical K-major no-swizzle core-matrix layout (8 rows x 16 B), the
// descriptors say LBO = 128 B (next K chunk), SBO = 256 B (next 8-row group).  Two CTAs share an SM (256 TMEM columns
// each), so one CTA's staging + MMA round trip hides behind the other's epilogue.
#include "tc_common.cuh"

#ifndef DR_LUT_PAIR
#define DR_LUT_PAIR 1
#endif
#ifndef DR_LUT_PREFETCH_B
#define DR_LUT_PREFETCH_B 1
#endif
#define TC_ROWS 128      // queries per CTA tile = TMEM lanes
#define TC_THREADS 256   // two threads per query row: warps 0-3 take the first half of a stage's centroids, warps 4-7 the second
#ifndef TC_NQ
#define TC_NQ 32         // centroids per stage
#endif
#define TC_COLS (4 * TC_NQ)   // TMEM columns per CTA (4 subspaces x TC_NQ centroids)
#define TC_CTAS (512 / TC_COLS)   // CTAs per SM: together they own the SM's 512 TMEM columns

namespace {

struct LutTcArgs {
    const float *codebook; const float *Q; long long B; int D, M, ds;
    uint32_t *out32;      // [B][M/4][256] packed words (search_fast.cu's global layout)
    float *scale, *offset; // phase 2 derives them from lo / range_bits (pq.cu:lut_u8_finalize_kernel's formulas) and stores them
    float *lo;            // [B][M] per-subspace minima (phase 1 writes, phase 2 reads)
    unsigned *range_bits; // [B] max range as float bits (phase 1, atomicMax)
    int words_per_cta;
};

// One stage: operands are in shared memory; thread 0 issues 4 x (nks + 1) MMAs and commits; stage_wait: everybody waits.
// (Measured and dropped: fetching the next stage's B operand into a second buffer between issue and wait — the staging warps then
// reach the epilogue late and the step gets 5 % slower; the other CTAs of the SM already cover the codebook loads.)
__device__ __forceinline__ void stage_mma(uint32_t tmem, uint32_t a_base, uint32_t b_base, int nks1, uint64_t *bar, uint32_t &phase,
                                          int tid) {
    fence_proxy_async();          // this thread's generic-proxy operand writes -> visible to the tensor core (async proxy)
    tc_fence_before();            // this thread's TMEM reads of the previous stage are ordered before the barrier
    __syncthreads();
    if (tid == 0) {
        tc_fence_after();
        constexpr uint32_t idesc = umma_idesc_tf32(TC_ROWS, TC_NQ);
        for (int j = 0; j < 4; ++j)
            for (int ks = 0; ks < nks1; ++ks)
                umma_tf32(tmem + (uint32_t)(j * TC_NQ), umma_desc(a_base + (uint32_t)((j * nks1 + ks) * (TC_ROWS * 32))),
                          umma_desc(b_base + (uint32_t)((j * nks1 + ks) * (TC_NQ * 32))), idesc, ks > 0 ? 1u : 0u);
        umma_commit(bar);
    }
}
__device__ __forceinline__ void stage_wait(uint64_t *bar, uint32_t &phase) {
    mbar_wait(bar, phase);
    phase ^= 1u;
    tc_fence_after();
}

// Stage the B operand of one stage: centroids [c0, c0 + TC_NQ) of the word's four subspaces (warp j stages subspace
// 4 w + j), rounded to TF32, plus the extra K-step [cn_hi, cn_lo, cn_hi, 1 | 1, 0, 0, 0] with cn = ||c||^2 in fp32.
__device__ __forceinline__ void stage_B(const float *__restrict__ codebook, int w, int c0, int ds, int nks, unsigned char *sB,
                                        int wid, int lane) {
    const int nks1 = nks + 1, j = wid;
    for (int rr = 0; rr < TC_NQ / 32; ++rr) {
        const int r = lane * (TC_NQ / 32) + rr;
        const float *src = codebook + ((size_t)(4 * w + j) * 256 + c0 + r) * ds;
        float cn = 0.0f;
        for (int kc = 0; kc < (ds >> 2); ++kc) {
            const float4 v = ldg_f4(src + kc * 4);
            cn = __fmaf_rn(v.x, v.x, cn); cn = __fmaf_rn(v.y, v.y, cn); cn = __fmaf_rn(v.z, v.z, cn); cn = __fmaf_rn(v.w, v.w, cn);
            *reinterpret_cast<float4 *>(sB + (j * nks1 + (kc >> 1)) * (TC_NQ * 32) + core_off(r, kc & 1)) = tf32_rna4(v);
        }
        const float ch = tf32_hi(cn);
        unsigned char *x = sB + (j * nks1 + nks) * (TC_NQ * 32);
        *reinterpret_cast<float4 *>(x + core_off(r, 0)) = make_float4(ch, cn - ch, ch, 1.0f);
        *reinterpret_cast<float4 *>(x + core_off(r, 1)) = make_float4(1.0f, 0.0f, 0.0f, 0.0f);
    }
}

// The same staging in two halves for the compile-time sub-dimensions: the loads of stage s+1 are issued right after the MMAs of
// stage s (their L2 round trip hides behind the MMA wait and the epilogue), the convert + store runs when the epilogue is done.
template <int DS>
__device__ __forceinline__ void stage_B_load(const float *__restrict__ codebook, int w, int c0, int wid, int lane, float4 (&v)[DS / 4]) {
    const float *src = codebook + ((size_t)(4 * w + wid) * 256 + c0 + lane) * DS;
#pragma unroll
    for (int kc = 0; kc < DS / 4; ++kc) v[kc] = ldg_f4(src + kc * 4);
}
template <int DS>
__device__ __forceinline__ void stage_B_store(const float4 (&v)[DS / 4], unsigned char *sB, int wid, int lane) {
    constexpr int nks = DS >> 3, nks1 = nks + 1;
    const int j = wid, r = lane;
    float cn = 0.0f;
#pragma unroll
    for (int kc = 0; kc < DS / 4; ++kc) {
        cn = __fmaf_rn(v[kc].x, v[kc].x, cn); cn = __fmaf_rn(v[kc].y, v[kc].y, cn);
        cn = __fmaf_rn(v[kc].z, v[kc].z, cn); cn = __fmaf_rn(v[kc].w, v[kc].w, cn);
        *reinterpret_cast<float4 *>(sB + (j * nks1 + (kc >> 1)) * (TC_NQ * 32) + core_off(r, kc & 1)) = tf32_rna4(v[kc]);
    }
    const float ch = tf32_hi(cn);
    unsigned char *x = sB + (j * nks1 + nks) * (TC_NQ * 32);
    *reinterpret_cast<float4 *>(x + core_off(r, 0)) = make_float4(ch, cn - ch, ch, 1.0f);
    *reinterpret_cast<float4 *>(x + core_off(r, 1)) = make_float4(1.0f, 0.0f, 0.0f, 0.0f);
}

// grid (query tiles, word groups), 256 threads: two per query row (= TMEM lane), each reducing half of a stage's centroids, so that
// eight warps per CTA (32 per SM) cover the TMEM round trips and the store latency.
// PHASE 1: lo[b][m] = min_c t and range_bits[b] = max(range_bits[b], max_c t - lo)        (then lut_u8_finalize_kernel)
// PHASE 2: out32[b][w][c] = the four subspaces' quantised entries, packed
// Dynamic shared memory: A blocks 4 * (nks+1) * 4 KB, then B blocks 4 * (nks+1) * TC_NQ * 32 B.
template <int PHASE, int DS>   // DS: compile-time sub-dimension (8 / 16 / 24), 0 = runtime
__global__ void __launch_bounds__(TC_THREADS, TC_CTAS) lut_u8_tc_kernel(const LutTcArgs a) {
    extern __shared__ __align__(1024) unsigned char tc_smem[];
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ uint32_t s_tmem;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int ds = DS ? DS : a.ds, nks = ds >> 3, nks1 = nks + 1, words = a.M >> 2, D = a.D, M = a.M;
    unsigned char *sA = tc_smem;
    unsigned char *sB = tc_smem + 4 * nks1 * (TC_ROWS * 32);
    float *sS = reinterpret_cast<float *>(sB + 4 * nks1 * (TC_NQ * 32));           // phase 1: [128 rows][8] lo / hi of the upper half
    const uint32_t a_base = smem_u32(sA), b_base = smem_u32(sB);

    if (tid == 0) { mbar_init(&s_bar, 1); fence_mbar_init(); }
    if (wid == 0) tmem_alloc(&s_tmem, TC_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem;
    const int half = wid >> 2, row = tid & (TC_ROWS - 1);
    const bool owner = tid < TC_ROWS;                                   // stages the A operand of its row, owns its statistics
    constexpr int HQ = TC_NQ / 2;                                       // centroids of a stage per thread (per subspace)
    const uint32_t tlane = tmem + ((uint32_t)((wid & 3) * 32) << 16) + (uint32_t)(half * HQ);   // this thread's lane quarter / column half
    uint32_t phase = 0;

    const long long b0 = (long long)blockIdx.x * TC_ROWS;
    const long long b = b0 + row;
    const bool live = b < a.B;
    const float *qrow = a.Q + (size_t)(live ? b : b0) * D;     // dead rows recompute row b0 and are never stored
    const int w_begin = blockIdx.y * a.words_per_cta;
    const int w_end = (w_begin + a.words_per_cta < words) ? (w_begin + a.words_per_cta) : words;

    float mul = -2.0f, inv = 1.0f, inv_hi = 1.0f, inv_lo = 0.0f;
    if (PHASE == 2) {
        // scale[b] = max range / 255 (1 if 0): every CTA of the tile recomputes it from the statistics pass; the CTAs of the
        // first word group also publish scale[b] and offset[b] = sum_m lo[b][m] + ||q_b||^2 (same operation order as
        // lut_u8_finalize_kernel), so no separate finalize launch is needed
        const float range = __uint_as_float(a.range_bits[live ? b : b0]);
        const float scale = range > 0.0f ? __fdiv_rn(range, 255.0f) : 1.0f;
        inv = __fdiv_rn(1.0f, scale);
        inv_hi = tf32_hi(inv); inv_lo = inv - inv_hi; mul = -2.0f * inv;
        if (blockIdx.y == 0 && live && owner) {
            float acc = 0.0f, qn = 0.0f;
            for (int m = 0; m < M; ++m) acc = __fadd_rn(acc, a.lo[(size_t)b * M + m]);
            if ((D & 3) == 0) {       // same element order as the scalar loop, 16-byte loads
                for (int j = 0; j < D; j += 4) {
                    const float4 v = ldg_f4(qrow + j);
                    qn = __fmaf_rn(v.x, v.x, qn); qn = __fmaf_rn(v.y, v.y, qn); qn = __fmaf_rn(v.z, v.z, qn); qn = __fmaf_rn(v.w, v.w, qn);
                }
            } else {
                for (int j = 0; j < D; ++j) { const float v = qrow[j]; qn = __fmaf_rn(v, v, qn); }
            }
            a.scale[b] = scale;
            a.offset[b] = __fadd_rn(acc, qn);
        }
    }
    float rmax = 0.0f;
    constexpr bool PF = DR_LUT_PREFETCH_B && (DS == 8 || (DS == 16 && PHASE == 1)) && TC_NQ == 32;   // register prefetch of the next stage's centroids
    float4 pfv[PF ? DS / 4 : 1];
    if (PF && wid < 4 && w_begin < w_end) stage_B_load<(PF ? DS : 4)>(a.codebook, w_begin, 0, wid, lane, pfv);
    for (int w = w_begin; w < w_end; ++w) {
        // A for this word: the row's 4 * ds query elements times -2 (phase 2: -2 inv), then the extra K-step
        for (int j = 0; j < 4 && owner; ++j) {
            const float *src = qrow + (size_t)(4 * w + j) * ds;
            for (int kc = 0; kc < (ds >> 2); ++kc) {
                float4 v = ldg_f4(src + kc * 4);
                v.x *= mul; v.y *= mul; v.z *= mul; v.w *= mul;
                *reinterpret_cast<float4 *>(sA + (j * nks1 + (kc >> 1)) * (TC_ROWS * 32) + core_off(row, kc & 1)) = tf32_rna4(v);
            }
            unsigned char *x = sA + (j * nks1 + nks) * (TC_ROWS * 32);
            if (PHASE == 1) {
                *reinterpret_cast<float4 *>(x + core_off(row, 0)) = make_float4(1.0f, 1.0f, 0.0f, 0.0f);
                *reinterpret_cast<float4 *>(x + core_off(row, 1)) = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            } else {
                const float k0 = -__fmul_rn(a.lo[(size_t)(live ? b : b0) * M + 4 * w + j], inv);
                const float kh = tf32_hi(k0);
                *reinterpret_cast<float4 *>(x + core_off(row, 0)) = make_float4(inv_hi, inv_hi, inv_lo, kh);
                *reinterpret_cast<float4 *>(x + core_off(row, 1)) = make_float4(k0 - kh, 0.0f, 0.0f, 0.0f);
            }
        }
        float lo[4], hi[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) { lo[j] = __int_as_float(0x7f800000); hi[j] = -__int_as_float(0x7f800000); }
        for (int c0 = 0; c0 < 256; c0 += TC_NQ) {
            if (wid < 4) {
                if (PF) stage_B_store<(PF ? DS : 4)>(pfv, sB, wid, lane);
                else stage_B(a.codebook, w, c0, ds, nks, sB, wid, lane);
            }
            stage_mma(tmem, a_base, b_base, nks1, &s_bar, phase, tid);
            if (PF && wid < 4) {
                const bool last_c = c0 + TC_NQ >= 256;
                const int wn = last_c ? w + 1 : w, cn0 = last_c ? 0 : c0 + TC_NQ;
                if (wn < w_end) stage_B_load<(PF ? DS : 4)>(a.codebook, wn, cn0, wid, lane, pfv);
            }
            stage_wait(&s_bar, phase);
            if (PHASE == 1) {
                // the four subspaces' loads go out together and are waited for once (a tcgen05.ld round trip is ~230 cycles with
                // eight warps on the port: one wait per load made this pass a chain of eight of them per stage)
#pragma unroll
                for (int i8 = 0; i8 < HQ / 8; ++i8) {
                    float v0[8], v1[8], v2[8], v3[8];
                    tmem_ld8(tlane + (uint32_t)(0 * TC_NQ + i8 * 8), v0);
                    tmem_ld8(tlane + (uint32_t)(1 * TC_NQ + i8 * 8), v1);
                    tmem_ld8(tlane + (uint32_t)(2 * TC_NQ + i8 * 8), v2);
                    tmem_ld8(tlane + (uint32_t)(3 * TC_NQ + i8 * 8), v3);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        lo[0] = fminf(lo[0], v0[i]); hi[0] = fmaxf(hi[0], v0[i]);
                        lo[1] = fminf(lo[1], v1[i]); hi[1] = fmaxf(hi[1], v1[i]);
                        lo[2] = fminf(lo[2], v2[i]); hi[2] = fmaxf(hi[2], v2[i]);
                        lo[3] = fminf(lo[3], v3[i]); hi[3] = fmaxf(hi[3], v3[i]);
                    }
                }
            } else {
#pragma unroll
                for (int i8 = 0; i8 < HQ / 8; ++i8) {
                    float v0[8], v1[8], v2[8], v3[8];
                    tmem_ld8(tlane + (uint32_t)(0 * TC_NQ + i8 * 8), v0);
                    tmem_ld8(tlane + (uint32_t)(1 * TC_NQ + i8 * 8), v1);
                    tmem_ld8(tlane + (uint32_t)(2 * TC_NQ + i8 * 8), v2);
                    tmem_ld8(tlane + (uint32_t)(3 * TC_NQ + i8 * 8), v3);
                    tmem_ld_wait();
                    uint32_t pk[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        uint32_t q0, q1, q2, q3;
                        asm("cvt.rni.sat.u8.f32 %0, %1;" : "=r"(q0) : "f"(v0[i]));
                        asm("cvt.rni.sat.u8.f32 %0, %1;" : "=r"(q1) : "f"(v1[i]));
                        asm("cvt.rni.sat.u8.f32 %0, %1;" : "=r"(q2) : "f"(v2[i]));
                        asm("cvt.rni.sat.u8.f32 %0, %1;" : "=r"(q3) : "f"(v3[i]));
                        pk[i] = q0 | (q1 << 8) | (q2 << 16) | (q3 << 24);
                    }
                    // A lane holds 32 contiguous bytes (one sector) of ITS row.  Lane pairs trade halves so that one store instruction
                    // writes whole sectors: first the even lane's row (even lane: words 0-3, odd lane: words 4-7), then the odd lane's.
                    const bool odd = lane & 1;
                    uint32_t s0[4], s1[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        s0[i] = __shfl_xor_sync(DR_FULL, odd ? pk[i] : pk[4 + i], 1);   // even gives A[4..7], odd gives B[0..3]
                    }
                    const long long br = b0 + (row & ~1);                                // the even lane's row; the odd lane's is br + 1
                    uint32_t *o = a.out32 + ((size_t)br * words + w) * 256 + c0 + half * HQ + i8 * 8;
                    const size_t rstride = (size_t)words * 256;
                    // instruction 1: row br      <- even lane: its own words 0-3 at +0; odd lane: the even lane's words 4-7 at +4
                    // instruction 2: row br + 1  <- even lane: the odd lane's words 0-3 at +0; odd lane: its own words 4-7 at +4
                    if (br < a.B)
                        *reinterpret_cast<uint4 *>(o + (odd ? 4 : 0)) = odd ? make_uint4(s0[0], s0[1], s0[2], s0[3]) : make_uint4(pk[0], pk[1], pk[2], pk[3]);
                    if (br + 1 < a.B)
                        *reinterpret_cast<uint4 *>(o + rstride + (odd ? 4 : 0)) = odd ? make_uint4(pk[4], pk[5], pk[6], pk[7]) : make_uint4(s0[0], s0[1], s0[2], s0[3]);
                    (void)s1;
                }
            }
        }
        if (PHASE == 1) {
            // the upper half's minima / maxima of the word join the lower half's through shared memory
            __syncthreads();
            if (!owner) {
#pragma unroll
                for (int j = 0; j < 4; ++j) { sS[row * 8 + j] = lo[j]; sS[row * 8 + 4 + j] = hi[j]; }
            }
            __syncthreads();
            if (owner) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float l = fminf(lo[j], sS[row * 8 + j]), h = fmaxf(hi[j], sS[row * 8 + 4 + j]);
                    rmax = fmaxf(rmax, h - l);
                    if (live) a.lo[(size_t)b * M + 4 * w + j] = l;
                }
            }
        }
    }
    if (PHASE == 1 && live && owner) atomicMax(a.range_bits + b, __float_as_uint(rmax));   // a non-negative float orders like its bits
    tc_fence_before();
    __syncthreads();
    if (wid == 0) tmem_dealloc(tmem, TC_COLS);
}

// ---- pair stages: N = 64 centroids per MMA ---------------------------------------------------------------------------------
// Same contraction, same 128 TMEM columns per CTA (four CTAs per SM), but a stage is TWO subspaces x 64 centroids instead of four
// x 32: every tcgen05.mma re-reads its 4 KB A tile (128 queries x 8) from shared memory whatever N is, so N = 64 halves the
// A-operand bytes per table entry (the kernel's shared-memory traffic is what bounds it).  A packed output word needs
// the bytes of all four subspaces of the code word: the stage of subspaces 0-1 leaves its 16-bit halves in registers (two
// centroids per register), the stage of subspaces 2-3 of the same 64 centroids completes the words and stores them.
#define T2_NQ 64
#define T2_COLS (2 * T2_NQ)
#define T2_CTAS (512 / T2_COLS)

// rows of the B operand of a pair stage: staging warp wid holds subspace jj = wid >> 1 of the pair, centroids (wid & 1) * 32 + lane
template <int DS>
__device__ __forceinline__ void stage_B2_load(const float *__restrict__ codebook, int w, int pr, int c0, int wid, int lane, float4 (&v)[DS / 4]) {
    const float *src = codebook + ((size_t)(4 * w + 2 * pr + (wid >> 1)) * 256 + c0 + (wid & 1) * 32 + lane) * DS;
#pragma unroll
    for (int kc = 0; kc < DS / 4; ++kc) v[kc] = ldg_f4(src + kc * 4);
}
template <int DS>
__device__ __forceinline__ void stage_B2_store(const float4 (&v)[DS / 4], unsigned char *sB, int wid, int lane) {
    constexpr int nks = DS >> 3, nks1 = nks + 1;
    const int jj = wid >> 1, r = (wid & 1) * 32 + lane;
    float cn = 0.0f;
#pragma unroll
    for (int kc = 0; kc < DS / 4; ++kc) {
        cn = __fmaf_rn(v[kc].x, v[kc].x, cn); cn = __fmaf_rn(v[kc].y, v[kc].y, cn);
        cn = __fmaf_rn(v[kc].z, v[kc].z, cn); cn = __fmaf_rn(v[kc].w, v[kc].w, cn);
        *reinterpret_cast<float4 *>(sB + (jj * nks1 + (kc >> 1)) * (T2_NQ * 32) + core_off(r, kc & 1)) = tf32_rna4(v[kc]);
    }
    const float ch = tf32_hi(cn);
    unsigned char *x = sB + (jj * nks1 + nks) * (T2_NQ * 32);
    *reinterpret_cast<float4 *>(x + core_off(r, 0)) = make_float4(ch, cn - ch, ch, 1.0f);
    *reinterpret_cast<float4 *>(x + core_off(r, 1)) = make_float4(1.0f, 0.0f, 0.0f, 0.0f);
}

template <int PHASE, int DS>   // DS = 8 (compile-time; other sub-dimensions use lut_u8_tc_kernel)
__global__ void __launch_bounds__(TC_THREADS, T2_CTAS) lut_u8_tc2_kernel(const LutTcArgs a) {
    extern __shared__ __align__(1024) unsigned char tc_smem[];
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ uint32_t s_tmem;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    constexpr int ds = DS, nks = DS >> 3, nks1 = nks + 1;
    const int words = a.M >> 2, D = a.D, M = a.M;
    unsigned char *sA = tc_smem;
    unsigned char *sB = tc_smem + 4 * nks1 * (TC_ROWS * 32);
    float *sS = reinterpret_cast<float *>(sB + 2 * nks1 * (T2_NQ * 32));
    const uint32_t a_base = smem_u32(sA), b_base = smem_u32(sB);

    if (tid == 0) { mbar_init(&s_bar, 1); fence_mbar_init(); }
    if (wid == 0) tmem_alloc(&s_tmem, T2_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem;
    const int half = wid >> 2, row = tid & (TC_ROWS - 1);
    const bool owner = tid < TC_ROWS;
    constexpr int HQ = T2_NQ / 2;                                       // 32 centroids of a stage per thread (per subspace)
    const uint32_t tlane = tmem + ((uint32_t)((wid & 3) * 32) << 16) + (uint32_t)(half * HQ);
    uint32_t phase = 0;

    const long long b0 = (long long)blockIdx.x * TC_ROWS;
    const long long b = b0 + row;
    const bool live = b < a.B;
    const float *qrow = a.Q + (size_t)(live ? b : b0) * D;
    const int w_begin = blockIdx.y * a.words_per_cta;
    const int w_end = (w_begin + a.words_per_cta < words) ? (w_begin + a.words_per_cta) : words;

    float mul = -2.0f, inv = 1.0f, inv_hi = 1.0f, inv_lo = 0.0f;
    if (PHASE == 2) {     // as in lut_u8_tc_kernel: scale from the statistics pass; the first word group publishes scale / offset
        const float range = __uint_as_float(a.range_bits[live ? b : b0]);
        const float scale = range > 0.0f ? __fdiv_rn(range, 255.0f) : 1.0f;
        inv = __fdiv_rn(1.0f, scale);
        inv_hi = tf32_hi(inv); inv_lo = inv - inv_hi; mul = -2.0f * inv;
        if (blockIdx.y == 0 && live && owner) {
            float acc = 0.0f, qn = 0.0f;
            for (int m = 0; m < M; ++m) acc = __fadd_rn(acc, a.lo[(size_t)b * M + m]);
            for (int j = 0; j < D; j += 4) {
                const float4 v = ldg_f4(qrow + j);
                qn = __fmaf_rn(v.x, v.x, qn); qn = __fmaf_rn(v.y, v.y, qn); qn = __fmaf_rn(v.z, v.z, qn); qn = __fmaf_rn(v.w, v.w, qn);
            }
            a.scale[b] = scale;
            a.offset[b] = __fadd_rn(acc, qn);
        }
    }
    float rmax = 0.0f;
    float4 pfv[DS / 4];
    if (wid < 4 && w_begin < w_end) stage_B2_load<DS>(a.codebook, w_begin, 0, 0, wid, lane, pfv);
    for (int w = w_begin; w < w_end; ++w) {
        for (int j = 0; j < 4 && owner; ++j) {
            const float *src = qrow + (size_t)(4 * w + j) * ds;
#pragma unroll
            for (int kc = 0; kc < (ds >> 2); ++kc) {
                float4 v = ldg_f4(src + kc * 4);
                v.x *= mul; v.y *= mul; v.z *= mul; v.w *= mul;
                *reinterpret_cast<float4 *>(sA + (j * nks1 + (kc >> 1)) * (TC_ROWS * 32) + core_off(row, kc & 1)) = tf32_rna4(v);
            }
            unsigned char *x = sA + (j * nks1 + nks) * (TC_ROWS * 32);
            if (PHASE == 1) {
                *reinterpret_cast<float4 *>(x + core_off(row, 0)) = make_float4(1.0f, 1.0f, 0.0f, 0.0f);
                *reinterpret_cast<float4 *>(x + core_off(row, 1)) = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            } else {
                const float k0 = -__fmul_rn(a.lo[(size_t)(live ? b : b0) * M + 4 * w + j], inv);
                const float kh = tf32_hi(k0);
                *reinterpret_cast<float4 *>(x + core_off(row, 0)) = make_float4(inv_hi, inv_hi, inv_lo, kh);
                *reinterpret_cast<float4 *>(x + core_off(row, 1)) = make_float4(k0 - kh, 0.0f, 0.0f, 0.0f);
            }
        }
        float lo[4], hi[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) { lo[j] = __int_as_float(0x7f800000); hi[j] = -__int_as_float(0x7f800000); }
        for (int c0 = 0; c0 < 256; c0 += T2_NQ) {
            uint32_t held[HQ / 2];              // phase 2: bytes of subspaces 0-1, two centroids per register
#pragma unroll
            for (int pr = 0; pr < 2; ++pr) {
                if (wid < 4) stage_B2_store<DS>(pfv, sB, wid, lane);
                fence_proxy_async();
                tc_fence_before();
                __syncthreads();
                if (tid == 0) {
                    tc_fence_after();
                    constexpr uint32_t idesc = umma_idesc_tf32(TC_ROWS, T2_NQ);
#pragma unroll
                    for (int jj = 0; jj < 2; ++jj)
#pragma unroll
                        for (int ks = 0; ks < nks1; ++ks)
                            umma_tf32(tmem + (uint32_t)(jj * T2_NQ), umma_desc(a_base + (uint32_t)(((2 * pr + jj) * nks1 + ks) * (TC_ROWS * 32))),
                                      umma_desc(b_base + (uint32_t)((jj * nks1 + ks) * (T2_NQ * 32))), idesc, ks > 0 ? 1u : 0u);
                    umma_commit(&s_bar);
                }
                if (wid < 4) {                  // the next stage's centroid rows travel under the MMA wait and the read-out
                    const bool last_c = c0 + T2_NQ >= 256;
                    const int wn = (pr == 1 && last_c) ? w + 1 : w;
                    const int cn0 = pr == 0 ? c0 : (last_c ? 0 : c0 + T2_NQ);
                    if (wn < w_end) stage_B2_load<DS>(a.codebook, wn, pr ^ 1, cn0, wid, lane, pfv);
                }
                stage_wait(&s_bar, phase);
                if (PHASE == 1) {
#pragma unroll
                    for (int i8 = 0; i8 < HQ / 8; ++i8) {
                        float v0[8], v1[8];
                        tmem_ld8(tlane + (uint32_t)(0 * T2_NQ + i8 * 8), v0);
                        tmem_ld8(tlane + (uint32_t)(1 * T2_NQ + i8 * 8), v1);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            lo[2 * pr] = fminf(lo[2 * pr], v0[i]); hi[2 * pr] = fmaxf(hi[2 * pr], v0[i]);
                            lo[2 * pr + 1] = fminf(lo[2 * pr + 1], v1[i]); hi[2 * pr + 1] = fmaxf(hi[2 * pr + 1], v1[i]);
                        }
                    }
                } else {
#pragma unroll
                    for (int i8 = 0; i8 < HQ / 8; ++i8) {
                        float v0[8], v1[8];
                        tmem_ld8(tlane + (uint32_t)(0 * T2_NQ + i8 * 8), v0);
                        tmem_ld8(tlane + (uint32_t)(1 * T2_NQ + i8 * 8), v1);
                        tmem_ld_wait();
                        uint32_t pk[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            uint32_t q0, q1;
                            asm("cvt.rni.sat.u8.f32 %0, %1;" : "=r"(q0) : "f"(v0[i]));
                            asm("cvt.rni.sat.u8.f32 %0, %1;" : "=r"(q1) : "f"(v1[i]));
                            const uint32_t h16 = q0 | (q1 << 8);
                            if (pr == 0) {
                                if (i & 1) held[i8 * 4 + (i >> 1)] |= h16 << 16; else held[i8 * 4 + (i >> 1)] = h16;
                            } else {
                                pk[i] = __byte_perm(held[i8 * 4 + (i >> 1)], h16, (i & 1) ? 0x5432 : 0x5410);
                            }
                        }
                        if (pr == 1) {
                            // lane pairs trade halves so that one store instruction writes whole 32-byte sectors (as in lut_u8_tc_kernel)
                            const bool odd = lane & 1;
                            uint32_t s0[4];
#pragma unroll
                            for (int i = 0; i < 4; ++i) s0[i] = __shfl_xor_sync(DR_FULL, odd ? pk[i] : pk[4 + i], 1);
                            const long long br = b0 + (row & ~1);
                            uint32_t *o = a.out32 + ((size_t)br * words + w) * 256 + c0 + half * HQ + i8 * 8;
                            const size_t rstride = (size_t)words * 256;
                            if (br < a.B)
                                *reinterpret_cast<uint4 *>(o + (odd ? 4 : 0)) = odd ? make_uint4(s0[0], s0[1], s0[2], s0[3]) : make_uint4(pk[0], pk[1], pk[2], pk[3]);
                            if (br + 1 < a.B)
                                *reinterpret_cast<uint4 *>(o + rstride + (odd ? 4 : 0)) = odd ? make_uint4(pk[4], pk[5], pk[6], pk[7]) : make_uint4(s0[0], s0[1], s0[2], s0[3]);
                        }
                    }
                }
            }
        }
        if (PHASE == 1) {
            __syncthreads();
            if (!owner) {
#pragma unroll
                for (int j = 0; j < 4; ++j) { sS[row * 8 + j] = lo[j]; sS[row * 8 + 4 + j] = hi[j]; }
            }
            __syncthreads();
            if (owner) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float l = fminf(lo[j], sS[row * 8 + j]), h = fmaxf(hi[j], sS[row * 8 + 4 + j]);
                    rmax = fmaxf(rmax, h - l);
                    if (live) a.lo[(size_t)b * M + 4 * w + j] = l;
                }
            }
        }
    }
    if (PHASE == 1 && live && owner) atomicMax(a.range_bits + b, __float_as_uint(rmax));
    tc_fence_before();
    __syncthreads();
    if (wid == 0) tmem_dealloc(tmem, T2_COLS);
}

}  // namespace

// Same contract as launch_lut_build_u8(..., word_layout = 1): needs M % 4 == 0 and (D / M) % 8 == 0.
// d_mn f32[B][M] and d_range u32[B] are scratch.  Two launches, both on the tensor cores and parallel over (128-query tile,
// group of code words): statistics, then quantise (which also publishes the per-query scale / offset).
int launch_lut_u8_finalize(const float *d_lo, const unsigned *d_range, const float *d_Q, int64_t B, int D, int M, float *d_scale,
                           float *d_offset, cudaStream_t s);   // pq.cu

int launch_lut_build_u8_tc(const float *d_codebook, const float *d_Q, int64_t B, int D, int M, uint8_t *d_out8, float *d_scale,
                           float *d_offset, float *d_mn, unsigned *d_range, int sms, cudaStream_t s) {
    DR_CHECK(M > 0 && D % M == 0 && (M & 3) == 0 && ((D / M) & 7) == 0,
             "dr_lut_build(u8, tensor cores): needs M %% 4 == 0 and a sub-dimension that is a multiple of 8 (D=%d M=%d)", D, M);
    if (B == 0) return 0;
    const int ds = D / M, nks1 = ds / 8 + 1, words = M / 4;
    const int smem = 4 * nks1 * (TC_ROWS * 32) + 4 * nks1 * (TC_NQ * 32) + TC_ROWS * 8 * 4;
    DR_CHECK(smem <= 200 * 1024, "dr_lut_build(u8, tensor cores): sub-dimension %d too large", ds);   // fewer CTAs per SM when large
    void (*k1)(const LutTcArgs) = lut_u8_tc_kernel<1, 0>;
    void (*k2)(const LutTcArgs) = lut_u8_tc_kernel<2, 0>;
    switch (ds) {
#if DR_LUT_PAIR
        case 8: k1 = lut_u8_tc2_kernel<1, 8>; k2 = lut_u8_tc2_kernel<2, 8>; break;     // pair stages (N = 64): same smem / TMEM footprint
        case 16: k1 = lut_u8_tc_kernel<1, 16>; k2 = lut_u8_tc_kernel<2, 16>; break;
#else
        case 8: k1 = lut_u8_tc_kernel<1, 8>; k2 = lut_u8_tc_kernel<2, 8>; break;
        case 16: k1 = lut_u8_tc_kernel<1, 16>; k2 = lut_u8_tc_kernel<2, 16>; break;
#endif
        case 24: k1 = lut_u8_tc_kernel<1, 24>; k2 = lut_u8_tc_kernel<2, 24>; break;
        default: break;
    }
    DR_CUDA(cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    DR_CUDA(cudaFuncSetAttribute(k2, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    LutTcArgs a;
    a.codebook = d_codebook; a.Q = d_Q; a.B = B; a.D = D; a.M = M; a.ds = ds;
    a.out32 = reinterpret_cast<uint32_t *>(d_out8); a.scale = d_scale; a.offset = d_offset; a.lo = d_mn; a.range_bits = d_range;
    const long long tiles = (B + TC_ROWS - 1) / TC_ROWS;
    // enough CTAs for >= 4 waves of the whole chip, a word group no smaller than 2 words (the A operand is staged per word)
    int groups = (int)((4LL * TC_CTAS * sms + tiles - 1) / tiles);
    if (groups > words / 2) groups = words / 2;
    if (groups < 1) groups = 1;
    a.words_per_cta = (words + groups - 1) / groups;
    groups = (words + a.words_per_cta - 1) / a.words_per_cta;
    DR_CHECK(tiles <= 2147483647LL, "dr_lut_build(u8, tensor cores): batch too large");
    dim3 grid((unsigned)tiles, (unsigned)groups);
    DR_CUDA(cudaMemsetAsync(d_range, 0, (size_t)B * 4, s));
    k1<<<grid, TC_THREADS, smem, s>>>(a);
    DR_LAUNCHED();
    k2<<<grid, TC_THREADS, smem, s>>>(a);
    DR_LAUNCHED();
    return 0;
}
