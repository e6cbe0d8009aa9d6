/*
 * oracle.c — CPU restatement of the pydiskann search / build hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the checker, never the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
 * Nothing under diskrag_b200/ links, imports or calls it.
 *
 * Parity status: PINNED.  tests/test_oracle_vs_reference.py runs the functions below against the
 * real reference compiled into oracle/_ref (oracle/build_ref.py), live; tests/test_golden_oracle.py and tests/test_golden_config0.py run them against the
 * committed golden vectors in tests/golden/ that were produced by that real reference (tests/golden/make_golden.py).
 *
 * Every function cites the reference file:line (relative to /root/reference) that it restates.
 * Where the reference's arithmetic order is knowable (heapq, sequential fp32 ADC sum, numpy
 * pairwise reduction) it is reproduced exactly; where it is not (BLAS sdot inside np.linalg.norm,
 * -ffast-math vectorisation) a named summation "flavor" is used and the tolerance is written in
 * the tests.
 *
 * Build: gcc -O2 -fPIC -shared -fopenmp -ffp-contract=off oracle.c -o liboracle.so -lm
 *        (-ffp-contract=off: no silent FMA contraction; where an FMA is part of the GPU's
 *        canonical order it is written as fmaf explicitly.)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------------------------------ */
/* Distances                                                                                  */
/* ------------------------------------------------------------------------------------------ */

/* cython_utils.pyx:18-24  l2_distance_fast_cython — squared L2, fp32 accumulator, i = 0..n-1.
 * (The reference is compiled -ffast-math so its own order may be vectorised: 1e-5 rel tolerance,
 * the same the reference's own known-answer test uses, scripts/test_pydiskann_cython.sh:36-56.) */
float orc_l2sq_seq(const float *x, const float *y, int n) {
    float acc = 0.0f;
    for (int i = 0; i < n; ++i) {
        float d = x[i] - y[i];
        acc += d * d;
    }
    return acc;
}

/* cython_utils.pyx:53-70  cosine_similarity_cython — returns 1 - cos (a distance); 0.0 if a norm
 * is zero.  fp32 accumulators; the final division is done in double because np.sqrt receives a
 * Python float. */
double orc_cosine_dist(const float *x, const float *y, int n) {
    float dot = 0.0f, nx = 0.0f, ny = 0.0f;
    for (int i = 0; i < n; ++i) {
        dot += x[i] * y[i];
        nx += x[i] * x[i];
        ny += y[i] * y[i];
    }
    if (nx == 0.0f || ny == 0.0f) return 0.0;
    return 1.0 - ((double)dot / (sqrt((double)nx) * sqrt((double)ny)));
}

/* numpy's float32 pairwise summation (numpy/_core/src/umath/loops_utils.h.src, pairwise_sum),
 * which is what np.sum(diff*diff, axis=1) runs per row at fast_pq.py:316 and what
 * np.sum(diff*diff) runs at search_engine.py:379.  numpy is a third-party dependency of the
 * reference (requirements.txt: numpy>=1.24.0, unpinned; 2.3.5 in this image); the algorithm is
 * restated from its published source and pinned empirically in tests/test_oracle_vs_reference.py. */
static float np_pairwise_sum_f32(const float *a, long n, long stride) {
    if (n < 8) {
        float res = 0.0f;
        for (long i = 0; i < n; ++i) res += a[i * stride];
        return res;
    } else if (n <= 128) {
        float r[8];
        for (int j = 0; j < 8; ++j) r[j] = a[j * stride];
        long i;
        for (i = 8; i < n - (n % 8); i += 8)
            for (int j = 0; j < 8; ++j) r[j] += a[(i + j) * stride];
        float res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; ++i) res += a[i * stride];
        return res;
    } else {
        long n2 = n / 2;
        n2 -= n2 % 8;
        return np_pairwise_sum_f32(a, n2, stride) + np_pairwise_sum_f32(a + n2 * stride, n - n2, stride);
    }
}

float orc_np_sum_f32(const float *a, long n) { return np_pairwise_sum_f32(a, n, 1); }

/* search_engine.py:374-379  _compute_exact_distance: diff = v - q ; np.sum(diff*diff)  (fp32). */
float orc_l2sq_numpy(const float *v, const float *q, int n) {
    float *t = (float *)malloc(sizeof(float) * (size_t)n);
    for (int i = 0; i < n; ++i) {
        float d = v[i] - q[i];
        t[i] = d * d;
    }
    float r = np_pairwise_sum_f32(t, n, 1);
    free(t);
    return r;
}

/* The canonical GPU summation order for an exact fp32 L2^2 (diskrag_b200/csrc/common.cuh,
 * warp_l2sq): 32 lanes; lane l owns elements (j*32+l)*VW + c (VW = 4 when n%4==0 else 1), each
 * lane accumulates with fmaf in increasing j,c; lanes are combined by an xor-butterfly
 * (offsets 16,8,4,2,1).  Restated here so that exact-distance traversals on the GPU can be
 * checked bit-for-bit; against the reference itself it is a tolerance comparison. */
float orc_l2sq_warp(const float *v, const float *q, int n) {
    float lane[32];
    int vw = (n % 4 == 0) ? 4 : 1;
    for (int l = 0; l < 32; ++l) {
        float acc = 0.0f;
        for (int base = l * vw; base < n; base += 32 * vw)
            for (int c = 0; c < vw; ++c) {
                float d = v[base + c] - q[base + c];
                acc = fmaf(d, d, acc);
            }
        lane[l] = acc;
    }
    for (int off = 16; off >= 1; off >>= 1) {
        float t[32];
        for (int l = 0; l < 32; ++l) t[l] = lane[l] + lane[l ^ off];
        memcpy(lane, t, sizeof(t));
    }
    return lane[0];
}

/* The GPU's canonical cosine distance (diskrag_b200/csrc/common.cuh:warp_cosdist, distance.cu:rowdist_kernel op 2):
 * the three sums in the lane / fmaf / xor-butterfly order of orc_l2sq_warp, the final division in double as
 * cosine_similarity_cython does (cython_utils.pyx:53-70); 0.0 when a norm is zero. */
static float warp_sum3(const float *a, const float *b, int n, int which) {
    float lane[32];
    int vw = (n % 4 == 0) ? 4 : 1;
    for (int l = 0; l < 32; ++l) {
        float acc = 0.0f;
        for (int base = l * vw; base < n; base += 32 * vw)
            for (int c = 0; c < vw; ++c) {
                float x = a[base + c], y = b[base + c];
                acc = which == 0 ? fmaf(x, y, acc) : (which == 1 ? fmaf(x, x, acc) : fmaf(y, y, acc));
            }
        lane[l] = acc;
    }
    for (int off = 16; off >= 1; off >>= 1) {
        float t[32];
        for (int l = 0; l < 32; ++l) t[l] = lane[l] + lane[l ^ off];
        memcpy(lane, t, sizeof(t));
    }
    return lane[0];
}
float orc_cosine_warp(const float *v, const float *q, int n) {
    float s0 = warp_sum3(v, q, n, 0), s1 = warp_sum3(v, q, n, 1), s2 = warp_sum3(v, q, n, 2);
    if (s1 == 0.0f || s2 == 0.0f) return 0.0f;
    return (float)(1.0 - ((double)s0 / (sqrt((double)s1) * sqrt((double)s2))));
}

/* flavor: 0 = double accumulation rounded once (stand-in for BLAS sdot inside np.linalg.norm,
 * vamana_graph.py:726,743), 1 = numpy pairwise, 2 = GPU warp order, 3 = sequential fp32. */
static float l2sq_flavor(const float *v, const float *q, int n, int flavor) {
    switch (flavor) {
    case 1: return orc_l2sq_numpy(v, q, n);
    case 2: return orc_l2sq_warp(v, q, n);
    case 3: return orc_l2sq_seq(v, q, n);
    default: {
        double acc = 0.0;
        for (int i = 0; i < n; ++i) {
            float d = v[i] - q[i];
            acc += (double)d * (double)d;
        }
        return (float)acc;
    }
    }
}
static float l2sq_refbuild(const float *x, const float *y, int n);
/* flavor 4 = the order g++ -O3 -ffast-math gives the reference's own `dist += (a-b)*(a-b)` loops (l2_distance_fast_cython and the
 * builder, see l2sq_refbuild below): bit-identical to the compiled reference in oracle/_ref. */
float orc_l2sq(const float *v, const float *q, int n, int flavor) { return flavor == 4 ? l2sq_refbuild(v, q, n) : l2sq_flavor(v, q, n, flavor); }

/* ------------------------------------------------------------------------------------------ */
/* Product quantiser                                                                          */
/* ------------------------------------------------------------------------------------------ */

/* fast_pq.py:294-318  DiskANNPQ.compute_distance_table — T[m,c] = sum_j (C[m,c,j]-q[m*ds+j])^2,
 * fp32, reduced per row by numpy's pairwise routine.  codebook is [M][256][ds] row-major. */
void orc_lut(const float *codebook, const float *q, int M, int ds, float *out) {
    float t[512];
    for (int m = 0; m < M; ++m)
        for (int c = 0; c < 256; ++c) {
            const float *cen = codebook + ((size_t)m * 256 + c) * ds;
            float *tt = ds <= 512 ? t : (float *)malloc(sizeof(float) * ds);
            for (int j = 0; j < ds; ++j) {
                float d = cen[j] - q[m * ds + j];
                tt[j] = d * d;
            }
            out[m * 256 + c] = np_pairwise_sum_f32(tt, ds, 1);
            if (tt != t) free(tt);
        }
}

/* Throughput-mode table (not a reference format; restated so the GPU's u8 mode can be checked bit-for-bit,
 * pq.cu: lut_u8_*): t[m,c] = ||c||^2 - 2 q_m.c as one fmaf chain (cn = fma chain of c_j^2, then fma(-2 q_j, c_j, .));
 * lo[m] = min_c t; range = max_m (max_c t - lo[m]); scale = range / 255 (1 if 0);
 * q[m,c] = clamp(rint(fma(t, inv, c0)), 0, 255) with inv = 1/scale, c0 = -(lo[m]*inv); offset = (sum_m lo[m], sequential) + ||q||^2 (fma chain).
 * The quantised ADC is the exact integer sum of q[m, code[m]]; d ~ offset + scale * sum. */
void orc_lut_u8(const float *codebook, const float *q, int M, int ds, uint8_t *out, float *scale_out, float *offset_out) {
    float *T = (float *)malloc(sizeof(float) * 256 * (size_t)M);
    float *mn = (float *)malloc(sizeof(float) * (size_t)M);
    float range = 0.0f, offset = 0.0f;
    for (int m = 0; m < M; ++m) {
        float lo = INFINITY, hi = -INFINITY;
        for (int c = 0; c < 256; ++c) {
            const float *cen = codebook + ((size_t)m * 256 + c) * ds;
            float acc = 0.0f;
            for (int j = 0; j < ds; ++j) acc = fmaf(cen[j], cen[j], acc);
            for (int j = 0; j < ds; ++j) acc = fmaf(-2.0f * q[m * ds + j], cen[j], acc);
            T[m * 256 + c] = acc;
            lo = fminf(lo, acc); hi = fmaxf(hi, acc);
        }
        mn[m] = lo;
        float r = hi - lo;
        if (r > range) range = r;
        offset += lo;
    }
    float qn = 0.0f;
    for (int j = 0; j < M * ds; ++j) qn = fmaf(q[j], q[j], qn);
    offset += qn;
    float scale = range > 0.0f ? range / 255.0f : 1.0f;
    const float inv = 1.0f / scale;
    for (int m = 0; m < M; ++m) {
        const float c0 = -(mn[m] * inv);
        for (int c = 0; c < 256; ++c) {
            float v = rintf(fmaf(T[m * 256 + c], inv, c0));
            out[m * 256 + c] = (uint8_t)fminf(fmaxf(v, 0.0f), 255.0f);
        }
    }
    *scale_out = scale; *offset_out = offset;
    free(T); free(mn);
}

/* fast_pq.py:320-328  asymmetric_distance_sq — acc = 0; for m in 0..M-1: acc += T[m, code[m]]
 * strictly sequential fp32. */
static inline float adc_seq(const uint8_t *code, const float *lut, int M) {
    float acc = 0.0f;
    for (int m = 0; m < M; ++m) acc += lut[m * 256 + code[m]];
    return acc;
}
/* Canonical GPU fast-mode order (search.cu, adc_tree): lane l sums m = l, l+32, ... sequentially,
 * then xor-butterfly 16..1.  Not a reference order: used to check the throughput mode bit-for-bit. */
static inline float adc_tree(const uint8_t *code, const float *lut, int M) {
    float lane[32];
    for (int l = 0; l < 32; ++l) {
        float acc = 0.0f;
        for (int m = l; m < M; m += 32) acc += lut[m * 256 + code[m]];
        lane[l] = acc;
    }
    for (int off = 16; off >= 1; off >>= 1) {
        float t[32];
        for (int l = 0; l < 32; ++l) t[l] = lane[l] + lane[l ^ off];
        memcpy(lane, t, sizeof(t));
    }
    return lane[0];
}
void orc_adc(const uint8_t *codes, const float *lut, long n, int M, int tree, float *out) {
    for (long i = 0; i < n; ++i)
        out[i] = tree ? adc_tree(codes + (size_t)i * M, lut, M) : adc_seq(codes + (size_t)i * M, lut, M);
}

/* fast_pq.py:245-267  encode — nearest centroid per subspace (sklearn KMeans.predict).  sklearn is a
 * third-party dependency (scikit-learn>=1.3.0 unpinned; 1.9.0 here) whose predict evaluates
 * ||c||^2 - 2 x.c in a chunked GEMM; here the published definition (argmin of the Euclidean
 * distance, lowest index on ties) is restated in double.  Agreement with sklearn is exact except
 * at provable near-ties (tests state the margin). */
void orc_pq_encode(const float *codebook, const float *X, long N, int D, int M, uint8_t *out) {
    int ds = D / M;
#pragma omp parallel for schedule(static)
    for (long i = 0; i < N; ++i)
        for (int m = 0; m < M; ++m) {
            const float *x = X + (size_t)i * D + m * ds;
            double best = INFINITY;
            int bi = 0;
            for (int c = 0; c < 256; ++c) {
                const float *cen = codebook + ((size_t)m * 256 + c) * ds;
                double acc = 0.0;
                for (int j = 0; j < ds; ++j) {
                    double d = (double)x[j] - (double)cen[j];
                    acc += d * d;
                }
                if (acc < best) { best = acc; bi = c; }
            }
            out[(size_t)i * M + m] = (uint8_t)bi;
        }
}

/* cython_utils.pyx:26-51  pq_distance_fast_cython — symmetric PQ distance, fp32 sequential. */
float orc_pq_sdc(const float *codebook, const uint8_t *c1, const uint8_t *c2, int M, int ds) {
    float total = 0.0f;
    for (int m = 0; m < M; ++m) {
        const float *a = codebook + ((size_t)m * 256 + c1[m]) * ds;
        const float *b = codebook + ((size_t)m * 256 + c2[m]) * ds;
        float s = 0.0f;
        for (int j = 0; j < ds; ++j) {
            float d = a[j] - b[j];
            s += d * d;
        }
        total += s;
    }
    return total;
}

/* ------------------------------------------------------------------------------------------ */
/* CPython heapq, restated exactly (Lib/heapq.py: heappush/_siftdown, heappop/_siftup) so that    */
/* even the array layout (which fixes the output order of exact ties after the stable sort)    */
/* matches the reference's heaps of (dist, id) / (-dist, id) tuples.                           */
/* ------------------------------------------------------------------------------------------ */
typedef struct { float d; int32_t id; } ent_t; /* tuple (d, id); for the result heap d is -dist */

static inline int ent_lt(ent_t a, ent_t b) { return a.d < b.d || (a.d == b.d && a.id < b.id); }

typedef struct { ent_t *a; int n, cap; } heap_t;

static void heap_init(heap_t *h, int cap) { h->a = (ent_t *)malloc(sizeof(ent_t) * (size_t)cap); h->n = 0; h->cap = cap; }
static void heap_free(heap_t *h) { free(h->a); }
static void heap_siftdown(heap_t *h, int startpos, int pos) {
    ent_t newitem = h->a[pos];
    while (pos > startpos) {
        int parentpos = (pos - 1) >> 1;
        ent_t parent = h->a[parentpos];
        if (ent_lt(newitem, parent)) { h->a[pos] = parent; pos = parentpos; continue; }
        break;
    }
    h->a[pos] = newitem;
}
static void heap_siftup(heap_t *h, int pos) {
    int endpos = h->n, startpos = pos;
    ent_t newitem = h->a[pos];
    int childpos = 2 * pos + 1;
    while (childpos < endpos) {
        int rightpos = childpos + 1;
        if (rightpos < endpos && !ent_lt(h->a[childpos], h->a[rightpos])) childpos = rightpos;
        h->a[pos] = h->a[childpos];
        pos = childpos;
        childpos = 2 * pos + 1;
    }
    h->a[pos] = newitem;
    heap_siftdown(h, startpos, pos);
}
static void heap_push(heap_t *h, ent_t e) {
    if (h->n == h->cap) { h->cap *= 2; h->a = (ent_t *)realloc(h->a, sizeof(ent_t) * (size_t)h->cap); }
    h->a[h->n++] = e;
    heap_siftdown(h, 0, h->n - 1);
}
static ent_t heap_pop(heap_t *h) {
    ent_t last = h->a[--h->n];
    if (h->n > 0) {
        ent_t ret = h->a[0];
        h->a[0] = last;
        heap_siftup(h, 0);
        return ret;
    }
    return last;
}
/* heapq.nsmallest(n, heap) + heapify: the resulting *set* is the n smallest tuples; the pop order
 * afterwards depends only on tuple order (ids are unique), so any heap layout is equivalent. */
static int ent_cmp(const void *pa, const void *pb) {
    ent_t a = *(const ent_t *)pa, b = *(const ent_t *)pb;
    return ent_lt(a, b) ? -1 : (ent_lt(b, a) ? 1 : 0);
}
static void heap_truncate_nsmallest(heap_t *h, int n) {
    if (h->n <= n) return;
    qsort(h->a, (size_t)h->n, sizeof(ent_t), ent_cmp); /* a sorted array is a valid heap */
    h->n = n;
}

/* visited set: byte map over N (the reference uses a Python set; membership is all that matters) */

/* stable sort of the result heap array by ascending distance == sorted(results, key=lambda x:-x[0])
 * (cython_utils.pyx:121-122; vamana_graph.py:640) */
static void stable_sort_by_dist(ent_t *a, int n) { /* insertion sort: stable, n <= L */
    for (int i = 1; i < n; ++i) {
        ent_t x = a[i];
        int j = i - 1;
        while (j >= 0 && a[j].d > x.d) { a[j + 1] = a[j]; --j; }
        a[j + 1] = x;
    }
}

/* ------------------------------------------------------------------------------------------ */
/* Variant A / B / D: best-first L-list search, literal two-heap form                          */
/*   A: cython_utils.pyx:72-122 greedy_search_cython (+ vamana_graph.py:301-329 ADC callback)   */
/*   B: vamana_graph.py:607-640 greedy_search (exact np.linalg.norm)                            */
/*   D: vamana_graph.py:719-760 beam_search_from_disk (exact, frontier truncated to beam_width) */
/* dist_mode: 0 = ADC sequential (needs codes+lut), 1 = sqrt(exact L2^2 flavor) as B/D,         */
/*            2 = exact L2^2 flavor (squared, as cython l2 callback), 3 = ADC tree (GPU fast),  */
/*            4 = u8 table, 5 = cosine distance 1 - cos (distance_metric='cosine').             */
/* truncate_frontier: D's `if len(beam) > beam_width: beam = nsmallest(beam_width, beam)`.      */
/* Outputs: out_ids/out_d = the whole result list in the reference's output order (<= L);       */
/*          trace (optional) = ids in the order their distance was computed (visited order).    */
/* Returns the length of the result list.                                                       */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
    const uint32_t *adj; int R; long N;
    const uint8_t *codes; int M; const float *lut; /* dist_mode 4: lut points at the u8 table */
    const float *vec; int D; const float *q; int flavor;
    int dist_mode;
} sctx_t;

static inline float node_dist(const sctx_t *c, long id) {
    switch (c->dist_mode) {
    case 0: return adc_seq(c->codes + (size_t)id * c->M, c->lut, c->M);
    case 3: return adc_tree(c->codes + (size_t)id * c->M, c->lut, c->M);
    case 4: { /* u8 table: exact integer sum, reported as a float (sums stay < 2^24) */
        const uint8_t *t8 = (const uint8_t *)c->lut, *code = c->codes + (size_t)id * c->M;
        uint32_t acc = 0;
        for (int m = 0; m < c->M; ++m) acc += t8[m * 256 + code[m]];
        return (float)acc;
    }
    case 1: return sqrtf(l2sq_flavor(c->vec + (size_t)id * c->D, c->q, c->D, c->flavor));
    case 5: /* distance_metric == 'cosine' (vamana_graph.py:325-326): flavor 2 = GPU warp order, else the reference's loop */
        return c->flavor == 2 ? orc_cosine_warp(c->vec + (size_t)id * c->D, c->q, c->D)
                              : (float)orc_cosine_dist(c->vec + (size_t)id * c->D, c->q, c->D);
    default: return orc_l2sq(c->vec + (size_t)id * c->D, c->q, c->D, c->flavor);   /* flavor 4: l2_distance_fast_cython's compiled order */
    }
}

/* deleted (may be NULL): is_deleted flags — lazily deleted nodes are never visited (cython_utils.pyx:100-109) and are filtered
 * from the output (:120); the caller resolves a deleted start (:84-90). */
static int search_heap_impl(const uint32_t *adj, int R, long N,
                    const uint8_t *codes, int M, const float *lut,
                    const float *vec, int D, const float *q, int flavor,
                    int dist_mode, int truncate_frontier,
                    int start, int L,
                    int32_t *out_ids, float *out_d, int32_t *out_hops, int32_t *out_nvisited,
                    int32_t *trace, int trace_cap, const uint8_t *deleted) {
    sctx_t c = {adj, R, N, codes, M, lut, vec, D, q, flavor, dist_mode};
    uint8_t *visited = (uint8_t *)calloc((size_t)N, 1);
    heap_t cand, res;
    heap_init(&cand, 4 * L + 64);
    heap_init(&res, L + 2);
    int hops = 0, nvis = 0;

    float d0 = node_dist(&c, start);
    visited[start] = 1;
    if (trace && nvis < trace_cap) trace[nvis] = start;
    ++nvis;
    heap_push(&cand, (ent_t){d0, start});
    heap_push(&res, (ent_t){-d0, start});

    while (cand.n > 0) {
        ent_t cur = heap_pop(&cand);
        if (deleted && deleted[cur.id]) continue;                   /* cython_utils.pyx:100-101 */
        float worst = -res.a[0].d;
        if (truncate_frontier) { /* D: `dist > worst and len(top_k) >= beam_width` (vamana_graph.py:734) */
            if (cur.d > worst && res.n >= L) break;
        } else {                 /* A/B: `dist > worst` (cython_utils.pyx:105) */
            if (cur.d > worst) break;
        }
        ++hops;
        const uint32_t *row = adj + (size_t)cur.id * R;
        for (int j = 0; j < R; ++j) {
            uint32_t nb = row[j];
            if ((long)nb >= N) continue; /* never true for a valid file; guards the byte map */
            if (visited[nb] || (deleted && deleted[nb])) continue;   /* :109 */
            visited[nb] = 1;
            float nd = node_dist(&c, nb);
            if (trace && nvis < trace_cap) trace[nvis] = (int32_t)nb;
            ++nvis;
            if (res.n < L || nd < -res.a[0].d) {
                heap_push(&cand, (ent_t){nd, (int32_t)nb});
                heap_push(&res, (ent_t){-nd, (int32_t)nb});
                if (res.n > L) heap_pop(&res);
            }
        }
        if (truncate_frontier) heap_truncate_nsmallest(&cand, L);
    }
    /* output order */
    int n = 0;
    ent_t *tmp = (ent_t *)malloc(sizeof(ent_t) * (size_t)(res.n > 0 ? res.n : 1));
    for (int i = 0; i < res.n; ++i)
        if (!(deleted && deleted[res.a[i].id])) tmp[n++] = (ent_t){-res.a[i].d, res.a[i].id};   /* :120 */
    if (truncate_frontier) qsort(tmp, (size_t)n, sizeof(ent_t), ent_cmp); /* D: sorted([(dist, id)]) full tuple order (vamana_graph.py:758) */
    else stable_sort_by_dist(tmp, n);                                      /* A/B: stable by dist */
    for (int i = 0; i < n; ++i) { out_ids[i] = tmp[i].id; out_d[i] = tmp[i].d; }
    free(tmp);
    if (out_hops) *out_hops = hops;
    if (out_nvisited) *out_nvisited = nvis;
    heap_free(&cand); heap_free(&res); free(visited);
    return n;
}

int orc_search_heap(const uint32_t *adj, int R, long N, const uint8_t *codes, int M, const float *lut,
                    const float *vec, int D, const float *q, int flavor, int dist_mode, int truncate_frontier,
                    int start, int L, int32_t *out_ids, float *out_d, int32_t *out_hops, int32_t *out_nvisited,
                    int32_t *trace, int trace_cap) {
    return search_heap_impl(adj, R, N, codes, M, lut, vec, D, q, flavor, dist_mode, truncate_frontier, start, L, out_ids, out_d,
                            out_hops, out_nvisited, trace, trace_cap, NULL);
}
int orc_search_heap_del(const uint32_t *adj, int R, long N, const uint8_t *codes, int M, const float *lut,
                        const float *vec, int D, const float *q, int flavor, int dist_mode, int truncate_frontier,
                        int start, int L, int32_t *out_ids, float *out_d, int32_t *out_hops, int32_t *out_nvisited,
                        int32_t *trace, int trace_cap, const uint8_t *deleted) {
    return search_heap_impl(adj, R, N, codes, M, lut, vec, D, q, flavor, dist_mode, truncate_frontier, start, L, out_ids, out_d,
                            out_hops, out_nvisited, trace, trace_cap, deleted);
}

/* ------------------------------------------------------------------------------------------ */
/* The "sorted L-list" formulation with W expansions per step — the exact statement of what the */
/* GPU kernel computes (search.cu).  W = 1 with strict_ties = 1 is provably equal to the heap   */
/* form above (ghost entries reproduce the one case where the heap form expands an evicted      */
/* node: an exact distance tie with the current worst) and is checked equal on every test graph.*/
/* W > 1 (DiskANN beam width) changes the visit order: throughput mode, recall-checked.         */
/*   key order: (dist, id) ascending.  Step: take the first W unexpanded list entries, in list  */
/*   order; for each, scan its adjacency row in stored order, first-seen ids are marked visited */
/*   and get a distance; all newcomers of the step are merged; the best L by key stay.          */
/* ------------------------------------------------------------------------------------------ */
typedef struct { float d; int32_t id; int expanded; } lent_t;

static inline int key_lt(float da, int32_t ia, float db, int32_t ib) { return da < db || (da == db && ia < ib); }

/* Throughput-mode option (DR "empty-step doubling", csrc/search_fast.cu): after a step none of whose newcomers entered the list, the
 * next step expands up to g_w_after_empty entries instead of W (0 = off).  Process-wide; set before a batch, read-only during it. */
static int g_w_after_empty = 0;
void orc_set_w_after_empty(int w2) { g_w_after_empty = w2; }
static int g_steps_total = 0, g_steps_empty = 0;   /* statistics of the last single-threaded calls (experiments) */
void orc_step_stats(int *total, int *empty, int reset) { *total = g_steps_total; *empty = g_steps_empty; if (reset) g_steps_total = g_steps_empty = 0; }

static int search_list_impl(const uint32_t *adj, int R, long N,
                    const uint8_t *codes, int M, const float *lut,
                    const float *vec, int D, const float *q, int flavor,
                    int dist_mode, int W, int strict_ties,
                    int start, int L,
                    int32_t *out_ids, float *out_d, int32_t *out_hops, int32_t *out_nvisited,
                    int32_t *trace, int trace_cap, const uint8_t *deleted) {
    sctx_t c = {adj, R, N, codes, M, lut, vec, D, q, flavor, dist_mode};
    uint8_t *visited = (uint8_t *)calloc((size_t)N, 1);
    lent_t *lst = (lent_t *)malloc(sizeof(lent_t) * (size_t)(L + 1));
    int gcap = 64, ng = 0;
    lent_t *ghost = (lent_t *)malloc(sizeof(lent_t) * (size_t)gcap); /* evicted, unexpanded, d == worst d */
    const int W2 = (!strict_ties && g_w_after_empty > W) ? g_w_after_empty : W;
    lent_t *nw = (lent_t *)malloc(sizeof(lent_t) * (size_t)(W2 * R + 1));
    int n = 0, hops = 0, nvis = 0;
    int Wcur = W;                   /* expansions allowed in the coming step */

    float d0 = node_dist(&c, start);
    visited[start] = 1;
    if (trace && nvis < trace_cap) trace[nvis] = start;
    ++nvis;
    lst[n++] = (lent_t){d0, start, 0};

    for (;;) {
        /* pick up to W unexpanded entries in key order (ghosts compete when strict_ties) */
        int picked[64], np_ = 0;
        int gi = -1;
        if (strict_ties && ng > 0) {
            if (ghost[0].d > lst[n - 1].d) ng = 0; /* worst improved: ghosts can no longer be popped before the break */
        }
        for (int i = 0; i < n && np_ < Wcur; ++i)
            if (!lst[i].expanded) picked[np_++] = i;
        if (strict_ties && ng > 0) {
            /* W == 1 here.  the ghost with the smallest id vs the first unexpanded list entry */
            int gbest = 0;
            for (int g = 1; g < ng; ++g) if (ghost[g].id < ghost[gbest].id) gbest = g;
            if (np_ == 0 || key_lt(ghost[gbest].d, ghost[gbest].id, lst[picked[0]].d, lst[picked[0]].id)) gi = gbest;
        }
        if (np_ == 0 && gi < 0) break;

        int nn = 0;
        int cur_ids[64];
        int ncur = 0;
        if (gi >= 0) {
            cur_ids[ncur++] = ghost[gi].id;
            ghost[gi] = ghost[--ng];
        } else {
            for (int i = 0; i < np_; ++i) { lst[picked[i]].expanded = 1; cur_ids[ncur++] = lst[picked[i]].id; }
        }
        for (int s = 0; s < ncur; ++s) {
            ++hops;
            const uint32_t *row = adj + (size_t)cur_ids[s] * R;
            for (int j = 0; j < R; ++j) {
                uint32_t nb = row[j];
                if ((long)nb >= N || visited[nb] || (deleted && deleted[nb])) continue;
                visited[nb] = 1;
                float nd = node_dist(&c, nb);
                if (trace && nvis < trace_cap) trace[nvis] = (int32_t)nb;
                ++nvis;
                nw[nn++] = (lent_t){nd, (int32_t)nb, 0};
            }
        }
        if (strict_ties) {
            /* sequential accept, strict '<' against the current worst, evict (max d, min id) */
            for (int t = 0; t < nn; ++t) {
                if (n >= L && !(nw[t].d < lst[n - 1].d)) continue;
                int pos = n;
                while (pos > 0 && key_lt(nw[t].d, nw[t].id, lst[pos - 1].d, lst[pos - 1].id)) { lst[pos] = lst[pos - 1]; --pos; }
                lst[pos] = nw[t];
                ++n;
                if (n > L) {
                    int e = n - 1; /* first entry of the last tie group = smallest id among the worst d */
                    while (e > 0 && lst[e - 1].d == lst[n - 1].d) --e;
                    lent_t ev = lst[e];
                    for (int i = e; i < n - 1; ++i) lst[i] = lst[i + 1];
                    --n;
                    if (!ev.expanded && ev.d == lst[n - 1].d) {
                        if (ng == gcap) { gcap *= 2; ghost = (lent_t *)realloc(ghost, sizeof(lent_t) * (size_t)gcap); }
                        ghost[ng++] = ev;
                    }
                }
            }
        } else {
            /* pure key-order merge: best L of old ∪ new */
            int entered = 0;
            const int full0 = n >= L;
            const lent_t worst0 = lst[n - 1];
            for (int t = 0; t < nn; ++t)
                if (!full0 || key_lt(nw[t].d, nw[t].id, worst0.d, worst0.id)) ++entered;   /* the kernel's survivor test (worst of the step's start) */
            Wcur = (entered == 0) ? W2 : W;
            __atomic_fetch_add(&g_steps_total, 1, __ATOMIC_RELAXED); __atomic_fetch_add(&g_steps_empty, entered == 0, __ATOMIC_RELAXED);
            for (int t = 0; t < nn; ++t) {
                if (n >= L && !key_lt(nw[t].d, nw[t].id, lst[n - 1].d, lst[n - 1].id)) continue;
                int pos = n;
                while (pos > 0 && key_lt(nw[t].d, nw[t].id, lst[pos - 1].d, lst[pos - 1].id)) { lst[pos] = lst[pos - 1]; --pos; }
                lst[pos] = nw[t];
                ++n;
                if (n > L) --n;
            }
        }
    }
    for (int i = 0; i < n; ++i) { out_ids[i] = lst[i].id; out_d[i] = lst[i].d; }
    if (out_hops) *out_hops = hops;
    if (out_nvisited) *out_nvisited = nvis;
    free(visited); free(lst); free(ghost); free(nw);
    return n;
}

int orc_search_list(const uint32_t *adj, int R, long N, const uint8_t *codes, int M, const float *lut,
                    const float *vec, int D, const float *q, int flavor, int dist_mode, int W, int strict_ties,
                    int start, int L, int32_t *out_ids, float *out_d, int32_t *out_hops, int32_t *out_nvisited,
                    int32_t *trace, int trace_cap) {
    return search_list_impl(adj, R, N, codes, M, lut, vec, D, q, flavor, dist_mode, W, strict_ties, start, L, out_ids, out_d,
                            out_hops, out_nvisited, trace, trace_cap, NULL);
}
int orc_search_list_del(const uint32_t *adj, int R, long N, const uint8_t *codes, int M, const float *lut,
                        const float *vec, int D, const float *q, int flavor, int dist_mode, int W, int strict_ties,
                        int start, int L, int32_t *out_ids, float *out_d, int32_t *out_hops, int32_t *out_nvisited,
                        int32_t *trace, int trace_cap, const uint8_t *deleted) {
    return search_list_impl(adj, R, N, codes, M, lut, vec, D, q, flavor, dist_mode, W, strict_ties, start, L, out_ids, out_d,
                            out_hops, out_nvisited, trace, trace_cap, deleted);
}

/* Rerank of a candidate list by exact squared L2: the composition the reference never writes as one
 * function: ids from the traversal, d2 as search_engine.py:374-379, stable sort by d2
 * (ties keep traversal order), first k.  flavor picks the fp32 summation order (1 numpy, 2 GPU). */
int orc_rerank(const float *vec, int D, const float *q, int flavor,
               const int32_t *ids, int n, int k, int32_t *out_ids, float *out_d) {
    ent_t *t = (ent_t *)malloc(sizeof(ent_t) * (size_t)(n > 0 ? n : 1));
    for (int i = 0; i < n; ++i) t[i] = (ent_t){l2sq_flavor(vec + (size_t)ids[i] * D, q, D, flavor), ids[i]};
    stable_sort_by_dist(t, n);
    int m = n < k ? n : k;
    for (int i = 0; i < m; ++i) { out_ids[i] = t[i].id; out_d[i] = t[i].d; }
    free(t);
    return m;
}

/* ------------------------------------------------------------------------------------------ */
/* Variant C: vamana_graph.py:535-605 beam_search_with_pq (and :690-717 beam_search, the same   */
/* loop without the delete checks), restated literally INCLUDING its inverted truncation:       */
/*   beam  = heapq min-heap of (dist, id); top_k = heapq min-heap of (-dist, id), capped at k;  */
/*   pop the best of beam; stop when it is worse than the worst of a full top_k (:580-581);     */
/*   every unvisited, undeleted neighbour gets a distance and enters BOTH heaps when top_k is    */
/*   not full or it beats top_k's worst (:589-593, evicting (max dist, min id));                */
/*   then `while len(beam) > beam_width: heappop(beam)` (:595-596) throws away the BEST entries */
/*   of the frontier and keeps the beam_width WORST ones.                                       */
/* Output (:599-601): the top_k heap ARRAY mapped to (sqrt(d), id) and sorted stably by         */
/* distance, so exact ties keep heap-array order; out_d holds sqrtf(d) when sqrt_out, else d.   */
/* deleted (may be NULL): is_deleted flags; the caller resolves a deleted start (:560-567).     */
/* Rows are scanned in stored order (the reference iterates a Python set; callers that want to  */
/* compare hand it lists in this order).  Returns the number of results (<= k).                 */
/* ------------------------------------------------------------------------------------------ */
int orc_beam_c(const uint32_t *adj, int R, long N,
               const uint8_t *codes, int M, const float *lut,
               const float *vec, int D, const float *q, int flavor,
               int dist_mode, const uint8_t *deleted, int sqrt_out,
               int start, int beam_width, int k,
               int32_t *out_ids, float *out_d, int32_t *out_hops, int32_t *out_nvisited,
               int32_t *trace, int trace_cap) {
    sctx_t c = {adj, R, N, codes, M, lut, vec, D, q, flavor, dist_mode};
    uint8_t *visited = (uint8_t *)calloc((size_t)N, 1);
    heap_t beam, top;
    heap_init(&beam, beam_width + R + 8);
    heap_init(&top, k + 2);
    int hops = 0, nvis = 0;

    float d0 = node_dist(&c, start);
    visited[start] = 1;
    if (trace && nvis < trace_cap) trace[nvis] = start;
    ++nvis;
    heap_push(&beam, (ent_t){d0, start});
    heap_push(&top, (ent_t){-d0, start});

    while (beam.n > 0) {
        ent_t cur = heap_pop(&beam);
        if (deleted && deleted[cur.id]) continue;                       /* :577-578 */
        if (cur.d > -top.a[0].d && top.n == k) break;                   /* :580-581 */
        ++hops;
        const uint32_t *row = adj + (size_t)cur.id * R;
        for (int j = 0; j < R; ++j) {
            uint32_t nb = row[j];
            if ((long)nb >= N) continue;
            if (visited[nb] || (deleted && deleted[nb])) continue;     /* :584 */
            visited[nb] = 1;
            float nd = node_dist(&c, nb);
            if (trace && nvis < trace_cap) trace[nvis] = (int32_t)nb;
            ++nvis;
            if (top.n < k || nd < -top.a[0].d) {                        /* :589 */
                heap_push(&beam, (ent_t){nd, (int32_t)nb});
                heap_push(&top, (ent_t){-nd, (int32_t)nb});
                if (top.n > k) heap_pop(&top);
            }
        }
        while (beam.n > beam_width) heap_pop(&beam);                    /* :595-596: drops the best */
    }
    int n = 0;
    ent_t *tmp = (ent_t *)malloc(sizeof(ent_t) * (size_t)(top.n > 0 ? top.n : 1));
    for (int i = 0; i < top.n; ++i) {
        if (deleted && deleted[top.a[i].id]) continue;                  /* :599 */
        float d = -top.a[i].d;
        tmp[n++] = (ent_t){sqrt_out ? sqrtf(d) : d, top.a[i].id};
    }
    stable_sort_by_dist(tmp, n);                                        /* :601 key = distance only */
    for (int i = 0; i < n; ++i) { out_ids[i] = tmp[i].id; out_d[i] = tmp[i].d; }
    free(tmp);
    if (out_hops) *out_hops = hops;
    if (out_nvisited) *out_nvisited = nvis;
    heap_free(&beam); heap_free(&top); free(visited);
    return n;
}

/* ------------------------------------------------------------------------------------------ */
/* Variant E: SearchEngineCorrect._pq_accelerated_graph_search (search_engine.py:398-506), the  */
/* served search, restated literally.  It is stochastic — a neighbour whose (non-squared!) ADC  */
/* distance lies in [0.8, 1.2) x the worst kept (squared) exact distance gets its exact distance */
/* with probability 0.2 (np.random.random() < 0.2, :394-395) — so the restatement carries       */
/* numpy's legacy generator: MT19937 seeded like np.random.seed(int) (init_genrand) and         */
/* random_sample() = genrand_res53.  With the same seed it reproduces the reference draw for    */
/* draw.  mt: 625 words of generator state (624 + position), owned by the caller.               */
/* ------------------------------------------------------------------------------------------ */
void orc_mt_seed(uint32_t *mt, uint32_t seed) {
    mt[0] = seed;
    for (int i = 1; i < 624; ++i) mt[i] = 1812433253u * (mt[i - 1] ^ (mt[i - 1] >> 30)) + (uint32_t)i;
    mt[624] = 624;
}
static uint32_t mt_next(uint32_t *mt) {
    if (mt[624] >= 624) {
        for (int i = 0; i < 624; ++i) {
            uint32_t y = (mt[i] & 0x80000000u) | (mt[(i + 1) % 624] & 0x7fffffffu);
            mt[i] = mt[(i + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        }
        mt[624] = 0;
    }
    uint32_t y = mt[mt[624]++];
    y ^= y >> 11; y ^= (y << 7) & 0x9d2c5680u; y ^= (y << 15) & 0xefc60000u; y ^= y >> 18;
    return y;
}
double orc_mt_random(uint32_t *mt) {
    uint32_t a = mt_next(mt) >> 5, b = mt_next(mt) >> 6;
    return ((double)a * 67108864.0 + (double)b) / 9007199254740992.0;
}

/* stats: [0] nodes_visited, [1] exact_distance_computations, [2] pq_distance_computations, [3] search_steps.
 * beam_width <= 0 = None (no frontier truncation, the /search default passes 8).  out_d = exact squared L2 in
 * numpy's summation order (np.sum(diff * diff), :374-379).  Returns min(k, len(results)). */
int orc_search_e(const uint32_t *adj, int R, long N,
                 const uint8_t *codes, int M, const float *lut,
                 const float *vec, int D, const float *q,
                 int start, int L, int k, int beam_width, uint32_t *mt,
                 int32_t *out_ids, float *out_d, int32_t *stats) {
    uint8_t *visited = (uint8_t *)calloc((size_t)N, 1);
    heap_t cand, res;
    heap_init(&cand, 4 * L + 64);
    heap_init(&res, L + 2);
    int nvis = 1, nexact = 0, npq = 0, steps = 0;
    const long max_steps = (long)L * 10 < N ? (long)L * 10 : N;                       /* :430 */
    visited[start] = 1;
    float d0 = orc_l2sq_numpy(vec + (size_t)start * D, q, D); ++nexact;
    heap_push(&cand, (ent_t){d0, start});
    heap_push(&res, (ent_t){-d0, start});
    while (cand.n > 0 && steps < max_steps) {
        ++steps;
        ent_t cur = heap_pop(&cand);
        if (res.n >= L && cur.d > -res.a[0].d) break;                                 /* :439-440 */
        const uint32_t *row = adj + (size_t)cur.id * R;
        for (int j = 0; j < R; ++j) {
            uint32_t nb = row[j];
            if ((long)nb >= N || visited[nb]) continue;
            visited[nb] = 1; ++nvis;
            float pq = sqrtf(adc_seq(codes + (size_t)nb * M, lut, M)); ++npq;       /* asymmetric_distance: sqrt (:368-369) */
            float worst = res.n ? -res.a[0].d : INFINITY;
            int take;                                                                 /* _should_compute_exact_distance (:381-397) */
            if (res.n < L) take = 1;
            else if (pq < worst * 0.8f) take = 1;                                     /* np.float32 * python float stays float32 */
            else if (pq < worst * 1.2f) take = orc_mt_random(mt) < 0.2;
            else take = 0;
            if (!take) continue;
            float ed = orc_l2sq_numpy(vec + (size_t)nb * D, q, D); ++nexact;
            if (res.n < L || ed < -res.a[0].d) {
                heap_push(&cand, (ent_t){ed, (int32_t)nb});
                heap_push(&res, (ent_t){-ed, (int32_t)nb});
                if (res.n > L) heap_pop(&res);
            }
        }
        if (beam_width > 0 && cand.n > beam_width) heap_truncate_nsmallest(&cand, beam_width);   /* :477-479 */
    }
    int n = res.n;
    ent_t *tmp = (ent_t *)malloc(sizeof(ent_t) * (size_t)(n > 0 ? n : 1));
    for (int i = 0; i < n; ++i) tmp[i] = (ent_t){-res.a[i].d, res.a[i].id};
    stable_sort_by_dist(tmp, n);                                                      /* :487 key = distance */
    int m = n < k ? n : k;
    for (int i = 0; i < m; ++i) { out_ids[i] = tmp[i].id; out_d[i] = tmp[i].d; }
    free(tmp);
    if (stats) { stats[0] = nvis; stats[1] = nexact; stats[2] = npq; stats[3] = steps; }
    heap_free(&cand); heap_free(&res); free(visited);
    return m;
}

/* Batched drivers for the CPU baseline (one query per OpenMP task).  form: 0 heap, 1 list. */
void orc_search_batch(const uint32_t *adj, int R, long N,
                      const uint8_t *codes, int M, const float *codebook,
                      const float *vec, int D, const float *Q, long B,
                      int dist_mode, int flavor, int W, int start, int L, int k, int rerank,
                      int32_t *out_ids, float *out_d, int32_t *out_hops, int32_t *out_nvisited, int nthreads) {
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel
    {
        float *lut = (float *)malloc(sizeof(float) * 256 * (size_t)(M > 0 ? M : 1));
        int32_t *ids = (int32_t *)malloc(sizeof(int32_t) * (size_t)(L + 1));
        float *ds_ = (float *)malloc(sizeof(float) * (size_t)(L + 1));
#pragma omp for schedule(dynamic, 4)
        for (long b = 0; b < B; ++b) {
            const float *q = Q + (size_t)b * D;
            float u8scale = 1.0f, u8off = 0.0f;
            if (dist_mode == 0 || dist_mode == 3) orc_lut(codebook, q, M, D / M, lut);
            if (dist_mode == 4) orc_lut_u8(codebook, q, M, D / M, (uint8_t *)lut, &u8scale, &u8off);
            int32_t h, v;
            int n;
            if (W <= 1)
                n = orc_search_heap(adj, R, N, codes, M, lut, vec, D, q, flavor, dist_mode, 0, start, L, ids, ds_, &h, &v, NULL, 0);
            else
                n = orc_search_list(adj, R, N, codes, M, lut, vec, D, q, flavor, dist_mode, W, 0, start, L, ids, ds_, &h, &v, NULL, 0);
            int m;
            if (rerank) m = orc_rerank(vec, D, q, flavor, ids, n, k, out_ids + b * k, out_d + b * k);
            else { m = n < k ? n : k; for (int i = 0; i < m; ++i) { out_ids[b * k + i] = ids[i]; out_d[b * k + i] = ds_[i]; } }
            for (int i = m; i < k; ++i) { out_ids[b * k + i] = -1; out_d[b * k + i] = INFINITY; }
            if (out_hops) out_hops[b] = h;
            if (out_nvisited) out_nvisited[b] = v;
        }
        free(lut); free(ids); free(ds_);
    }
}

/* ------------------------------------------------------------------------------------------ */
/* Vamana build, sequential, restated from cython_utils.pyx:269-492 including its quirks, so    */
/* that with the same permutations and medoid it can be compared with the real reference.       */
/* ------------------------------------------------------------------------------------------ */
typedef struct { float d; int id; } pfi_t; /* std::pair<float,int>, lexicographic */
static int pfi_cmp(const void *a, const void *b) {
    pfi_t x = *(const pfi_t *)a, y = *(const pfi_t *)b;
    if (x.d < y.d) return -1;
    if (x.d > y.d) return 1;
    return (x.id > y.id) - (x.id < y.id);
}
/* The distance loops INSIDE the C++ builder (cython_utils.pyx:384-386, 411-413, 452-454, 477-479).  The source is the same
 * sequential `dist += (a-b)*(a-b)`, but the reference is compiled -O3 -ffast-math (pydiskann/setup.py:10) and its own arithmetic is
 * what the compiler made of that: g++ 13 on x86-64 (SSE2 baseline) vectorises the strided memoryview loop two elements at a time
 * (objdump of oracle/_ref: movss/unpcklps pairs, subps, mulps, addps into ONE two-lane accumulator, then lane0 + lane1, then a
 * scalar tail; n <= 4 stays scalar).  I.e. even-indexed and odd-indexed terms are summed separately and added at the end.  The
 * sequential build is chaotic — ONE near-tie decided differently (measured: d(p*,p') vs d(p,p') 4 ulps apart at insertion 469 of a
 * 1500-point build) changes almost every row afterwards — so the builder restatement has to use the compiled order to stay
 * row-identical beyond a few hundred points.  Compiler-specific by nature: tests/test_oracle_vs_reference.py checks it live. */
static float l2sq_refbuild(const float *x, const float *y, int n) {
    if (n <= 4) return orc_l2sq_seq(x, y, n);
    float s0 = 0.0f, s1 = 0.0f;
    const int h = n >> 1;
    for (int i = 0; i < h; ++i) {
        float d0 = x[2 * i] - y[2 * i], d1 = x[2 * i + 1] - y[2 * i + 1];
        s0 += d0 * d0;
        s1 += d1 * d1;
    }
    float s = s0 + s1;
    if (n & 1) { float d = x[n - 1] - y[n - 1]; s += d * d; }
    return s;
}

typedef struct { int *v; int n, cap; } ivec_t;
static void iv_push(ivec_t *a, int x) {
    if (a->n == a->cap) { a->cap = a->cap ? a->cap * 2 : 8; a->v = (int *)realloc(a->v, sizeof(int) * (size_t)a->cap); }
    a->v[a->n++] = x;
}
static int int_cmp(const void *a, const void *b) { return (*(const int *)a > *(const int *)b) - (*(const int *)a < *(const int *)b); }

/* cython_utils.pyx:371-433  greedy_search_fast_cython.  `results` is a FIFO window, not a heap:
 * the accept test and the early exit look at results[0] (the OLDEST retained entry), eviction
 * drops the oldest (:400,416,423); `candidates` is a vector re-sorted after every expansion and
 * popped from the front (:397,426).  Returns the window ids in window order. */
static int build_search(const float *P, int D, const ivec_t *adj, long N, int start, int qi, int L,
                        int *out, uint32_t *vis_stamp, uint32_t stamp) {
    const float *q = P + (size_t)qi * D;
    pfi_t *cand = NULL; int nc = 0, ccap = 0;
    pfi_t *res = (pfi_t *)malloc(sizeof(pfi_t) * (size_t)(L + 2)); int nr = 0;
    float dist = l2sq_refbuild(P + (size_t)start * D, q, D);
#define CPUSH(e) do { if (nc == ccap) { ccap = ccap ? ccap * 2 : 256; cand = (pfi_t *)realloc(cand, sizeof(pfi_t) * (size_t)ccap); } cand[nc++] = (e); } while (0)
    CPUSH(((pfi_t){dist, start}));
    res[nr++] = (pfi_t){-dist, start};
    vis_stamp[start] = stamp;
    int head = 0; /* candidates.erase(begin) == advance head (the vector is fully re-sorted each round) */
    while (head < nc) {
        dist = cand[head].d; int cur = cand[head].id; ++head;
        if (nr >= L && dist > -res[0].d) break;
        for (int i = 0; i < adj[cur].n; ++i) {
            int nb = adj[cur].v[i];
            if (vis_stamp[nb] == stamp) continue;
            vis_stamp[nb] = stamp;
            float nd = l2sq_refbuild(P + (size_t)nb * D, q, D);
            if (nr < L || nd < -res[0].d) {
                /* compact the consumed prefix lazily so that push_back semantics hold */
                CPUSH(((pfi_t){nd, nb}));
                res[nr++] = (pfi_t){-nd, nb};
                if (nr > L) { memmove(res, res + 1, sizeof(pfi_t) * (size_t)(nr - 1)); --nr; }
            }
        }
        qsort(cand + head, (size_t)(nc - head), sizeof(pfi_t), pfi_cmp);
    }
#undef CPUSH
    for (int i = 0; i < nr; ++i) out[i] = res[i].id;
    free(cand); free(res);
    return nr;
}

/* cython_utils.pyx:435-492  robust_prune_fast_cython.  alpha multiplies SQUARED distances (:483).
 * The outer loop bound is evaluated once (Cython caches range(size()), see the generated C++), so
 * after erasures `i` may run into the stale tail of the vector; std::vector::erase leaves the old
 * bytes there, which is reproduced with a plain array + logical size.  The selected set is a
 * std::set<int>, so the row is written in ascending id order (:490-492). */
static void build_prune(const float *P, int D, ivec_t *adj, int p, const int *cset, int ncset, float alpha, int R) {
    pfi_t *cw = (pfi_t *)malloc(sizeof(pfi_t) * (size_t)(ncset + 1)); int n = 0;
    for (int t = 0; t < ncset; ++t) {
        int cid = cset[t];
        if (cid == p) continue;
        cw[n++] = (pfi_t){l2sq_refbuild(P + (size_t)p * D, P + (size_t)cid * D, D), cid};
    }
    qsort(cw, (size_t)n, sizeof(pfi_t), pfi_cmp);
    int bound = n; /* cached loop bound */
    int *sel = (int *)malloc(sizeof(int) * (size_t)(R + 1)); int ns = 0;
    for (int i = 0; i < bound; ++i) {
        if (ns >= R) break;
        int pstar = cw[i].id;
        int dup = 0;
        for (int s = 0; s < ns; ++s) if (sel[s] == pstar) { dup = 1; break; }
        if (!dup) sel[ns++] = pstar;
        int j = i + 1;
        while (j < n) {
            int pp = cw[j].id;
            int insel = 0;
            for (int s = 0; s < ns; ++s) if (sel[s] == pp) { insel = 1; break; }
            if (insel) { ++j; continue; }
            float dsp = l2sq_refbuild(P + (size_t)pstar * D, P + (size_t)pp * D, D);
            if (alpha * dsp <= cw[j].d) {
                memmove(cw + j, cw + j + 1, sizeof(pfi_t) * (size_t)(n - j - 1)); /* erase(begin+j); tail byte image stays */
                --n;
            } else ++j;
        }
    }
    qsort(sel, (size_t)ns, sizeof(int), int_cmp);
    adj[p].n = 0;
    for (int s = 0; s < ns; ++s) iv_push(&adj[p], sel[s]);
    free(cw); free(sel);
}

/* cython_utils.pyx:269-369  build_vamana_index_cython: two passes (alpha 1.0 then alpha) over the
 * given permutations (the reference draws them with Python's random.shuffle, :303-308), starting
 * from an EMPTY adjacency (:287).  out_adj is [N][cap] padded with -1, out_deg[N]. Rows can exceed R
 * transiently only inside the loop; final rows are <= R. */
int orc_vamana_build(const float *P, long N, int D, int R, int L, float alpha, int medoid,
                     const int32_t *sigma0, const int32_t *sigma1, int32_t *out_adj, int cap, int32_t *out_deg) {
    ivec_t *adj = (ivec_t *)calloc((size_t)N, sizeof(ivec_t));
    uint32_t *stampv = (uint32_t *)calloc((size_t)N, sizeof(uint32_t));
    uint32_t stamp = 0;
    int *win = (int *)malloc(sizeof(int) * (size_t)(L + 2));
    int *cset = (int *)malloc(sizeof(int) * (size_t)(L + 4 * R + 64));
    for (int pass = 0; pass < 2; ++pass) {
        const int32_t *sigma = pass == 0 ? sigma0 : sigma1;
        float a = pass == 0 ? 1.0f : alpha;
        for (long k = 0; k < N; ++k) {
            int idx = sigma[k];
            int nw = build_search(P, D, adj, N, medoid, idx, L, win, stampv, ++stamp);
            /* candidate_set = std::set(window ∪ N(idx)): ascending unique ids */
            int nc = 0;
            for (int i = 0; i < nw; ++i) cset[nc++] = win[i];
            for (int i = 0; i < adj[idx].n; ++i) cset[nc++] = adj[idx].v[i];
            qsort(cset, (size_t)nc, sizeof(int), int_cmp);
            int u = 0;
            for (int i = 0; i < nc; ++i) if (i == 0 || cset[i] != cset[i - 1]) cset[u++] = cset[i];
            build_prune(P, D, adj, idx, cset, u, a, R);
            /* reverse edges (:335-353); adj[idx] is not modified by pruning a different node */
            for (int i = 0; i < adj[idx].n; ++i) {
                int nb = adj[idx].v[i];
                if (nb == idx) continue;
                int exists = 0;
                for (int j = 0; j < adj[nb].n; ++j) if (adj[nb].v[j] == idx) { exists = 1; break; }
                if (!exists) iv_push(&adj[nb], idx);
                if (adj[nb].n > R) {
                    int m = adj[nb].n;
                    int *cs = (int *)malloc(sizeof(int) * (size_t)m);
                    memcpy(cs, adj[nb].v, sizeof(int) * (size_t)m);
                    qsort(cs, (size_t)m, sizeof(int), int_cmp);
                    int uu = 0;
                    for (int t = 0; t < m; ++t) if (t == 0 || cs[t] != cs[t - 1]) cs[uu++] = cs[t];
                    build_prune(P, D, adj, nb, cs, uu, a, R);
                    free(cs);
                }
            }
        }
    }
    int overflow = 0;
    for (long i = 0; i < N; ++i) {
        out_deg[i] = adj[i].n;
        for (int j = 0; j < cap; ++j) out_adj[(size_t)i * cap + j] = j < adj[i].n ? adj[i].v[j] : -1;
        if (adj[i].n > cap) overflow = 1;
        free(adj[i].v);
    }
    free(adj); free(stampv); free(win); free(cset);
    return overflow;
}

/* cython_utils.pyx:210-263  compute_approximate_medoid_cython, given the sample ids (the reference
 * draws them from a time-seeded mt19937): argmin over samples of sum_j ||x_s - x_j|| with an fp32
 * inner sum, sqrt, and an fp64 outer sum.  For N <= sample the j == i term is skipped (:228). */
int orc_medoid(const float *P, long N, int D, const int32_t *samples, int ns, int skip_self) {
    double best = INFINITY;
    int bi = 0;
    for (int s = 0; s < ns; ++s) {
        const float *x = P + (size_t)samples[s] * D;
        double sum = 0.0;
        for (long j = 0; j < N; ++j) {
            if (skip_self && j == samples[s]) continue;
            float d = orc_l2sq_seq(x, P + (size_t)j * D, D);
            sum += pow((double)d, 0.5);
        }
        if (sum < best) { best = sum; bi = s; }
    }
    return samples[bi];
}

/* Brute-force ground truth (dataset_benchmark.py:62-73 compute_ground_truth: argsort of
 * np.linalg.norm): top-k ids by exact distance, double accumulation, ties by id. */
void orc_ground_truth(const float *X, long N, int D, const float *Q, long B, int k, int32_t *out) {
#pragma omp parallel for schedule(dynamic, 1)
    for (long b = 0; b < B; ++b) {
        pfi_t *top = (pfi_t *)malloc(sizeof(pfi_t) * (size_t)(k + 1));
        int n = 0;
        for (long i = 0; i < N; ++i) {
            double acc = 0.0;
            const float *x = X + (size_t)i * D, *q = Q + (size_t)b * D;
            for (int j = 0; j < D; ++j) { double d = (double)x[j] - (double)q[j]; acc += d * d; }
            pfi_t e = {(float)acc, (int)i};
            if (n < k || pfi_cmp(&e, &top[n - 1]) < 0) {
                int pos = n < k ? n : k - 1;
                while (pos > 0 && pfi_cmp(&e, &top[pos - 1]) < 0) { top[pos] = top[pos - 1]; --pos; }
                top[pos] = e;
                if (n < k) ++n;
            }
        }
        for (int i = 0; i < k; ++i) out[b * k + i] = i < n ? top[i].id : -1;
        free(top);
    }
}

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
