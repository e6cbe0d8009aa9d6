/*
 * diskrag_b200.h — C ABI of the B200-native pydiskann hot path (libdiskrag_b200.so).
 *
 * The reference (Jolara-ai/diskrag) has no FFI registry: its only native unit is the Cython module
 * pydiskann/cython_utils.pyx and the boundary is Python call signatures.  Each entry
 * point below names the reference interface it replaces (file:line relative to the reference root).
 * The Python shims in diskrag_b200/ (vamana_graph.py, cython_utils.py, pq/fast_pq.py,
 * io/diskann_persist.py) bind these with ctypes and present the reference's names; INTEGRATION.md shows
 * the stub a maintainer adds on the reference side.
 *
 * Conventions
 *   - plain pointers and sizes only; no torch / numpy types.
 *   - every function returns 0 on success, non-zero on failure; dr_last_error() returns a
 *     thread-local message.  There is no CPU fallback: without a CUDA device every compute call fails.
 *   - "host" functions take host pointers and do the H2D / D2H copies themselves (the drop-in path,
 *     numpy-owned buffers, nothing retained after return).  "_dev" functions take device pointers
 *     on the index's device and a cudaStream_t (as void*), and only enqueue work.
 *   - arrays are C-contiguous, row-major.  ids are int32 (N < 2^31).
 *   - threading: the entry points that use a handle's scratch (search, table build, delete mask, export) hold a
 *     per-handle mutex for the duration of the call, so concurrent callers on one handle are serialised rather than
 *     racing; different handles are independent.  Work enqueued by the "_dev" entries of one handle must stay on one
 *     stream at a time (the next call reuses the handle's table scratch).
 */
#ifndef DISKRAG_B200_H
#define DISKRAG_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DR_ABI_VERSION 2

typedef struct dr_index dr_index; /* opaque device-resident index (vectors, adjacency, PQ codes, codebook) */

/* traversal distance */
#define DR_DIST_PQ 0    /* ADC over PQ codes: vamana_graph.py:301-329 compute_query_distance + fast_pq.py:320-328 */
#define DR_DIST_EXACT 1 /* fp32 squared L2 on the full vectors: vamana_graph.py:719-760, :607-640 */
#define DR_DIST_COSINE 2 /* exact traversal with 1 - cos(x, q) (distance_metric='cosine', vamana_graph.py:301-329 -> cython_utils.pyx:53-70) */
/* ADC summation order */
#define DR_ADC_SEQ 0  /* m = 0..M-1 sequential fp32 adds: bit-identical to fast_pq.py:320-328 */
#define DR_ADC_TREE 1 /* lane-strided partial sums + butterfly: throughput mode, not a reference order */

/* ADC table format held in shared memory */
#define DR_LUT_F32 0 /* M x 256 fp32, bit-identical to DiskANNPQ.compute_distance_table */
#define DR_LUT_U8 1  /* M x 256 bytes: q = rint((T - min_m) / scale), scale = max range / 255; exact integer sums */
#define DR_LUT_U8_TC 2 /* the same table built on the tensor cores (tcgen05, TF32 products): entries within one unit of DR_LUT_U8;
                          needs M % 4 == 0 and (D / M) % 8 == 0 */

/* per-query status bits written to out_status */
#define DR_ST_OK 0
#define DR_ST_VISITED_OVERFLOW 1 /* visited set exceeded smem + overflow table: result incomplete */
#define DR_ST_TIE_OVERFLOW 2     /* more than 16 exact-tie ghosts at the list boundary */

typedef struct dr_search_params {
    int32_t k;         /* results per query */
    int32_t L;         /* search list size (reference: L / beam_width), 1..512 */
    int32_t W;         /* nodes expanded per step; 1 = the reference's visit order (bit-exact mode) */
    int32_t dist;      /* DR_DIST_PQ | DR_DIST_EXACT */
    int32_t adc_order; /* DR_ADC_SEQ | DR_ADC_TREE (DR_DIST_PQ only) */
    int32_t rerank;    /* 1: exact fp32 L2^2 rerank of the final list (search_engine.py:374-379), stable */
    int32_t sqrt_out;  /* 1: report sqrt(d2) like np.linalg.norm (vamana_graph.py:726,743) */
    int32_t hash_cap;  /* 0 = auto; > 0 forces the shared-memory visited table size (power of two; tests);
                          < 0 (DR_LUT_U8*) keeps the visited set in a per-CTA global table (L2-resident) instead */
    int32_t chunk;     /* 0 = auto; queries per launch (bounds the device LUT buffer) */
    int32_t threads;   /* 0 = auto; CTA size (multiple of 32) */
    int32_t lut_fmt;   /* DR_LUT_F32 (reference arithmetic) | DR_LUT_U8 (throughput: 8-bit table, integer sums) */
    int32_t prefetch;  /* DR_LUT_U8 only, results unchanged; bit mask of L2 prefetches: 1 = the adjacency row of every accepted
                          candidate; 2 = speculative: the rows of the W entries next in line and of accepted candidates ranking
                          before them; 4 = the PQ code row of a neighbour as soon as it is first seen; 8 = the code rows of the
                          unseen neighbours of the W entries next in line (one step ahead of the step that needs them);
                          16 = not an L2 prefetch: the adjacency rows of the W selected entries are bulk-copied to shared memory
                          while the merge that selected them is still finishing (needs R % 4 == 0 and W <= 16, ignored otherwise) */
    int32_t start_plus1; /* entry point of THIS call (the reference passes start_idx per call): 0 = the index's own (medoid /
                          dr_index_set_start), s + 1 = node s.  Per call, so concurrent callers never see each other's start */
    int32_t w_after_empty; /* DR_LUT_U8* only.  0 / <= W: off.  Otherwise a step that follows a step WITHOUT survivors (no newcomer
                          entered the list, so the list and hence the next entries to expand are unchanged) expands up to this many
                          entries instead of W: the same expansions in fewer barrier-separated steps ("empty-step doubling").
                          Restated by the oracle (orc_set_w_after_empty); <= 32 */
    int32_t ignore_deleted; /* 1: this call does not look at the lazy-delete mask — greedy_search / greedy_search_optimized
                          (vamana_graph.py:607-640, 762-793) never test is_deleted; greedy_search_cython does (cython_utils.pyx:84-120) */
} dr_search_params;

/* ---- library ---- */
int dr_abi_version(void);
const char *dr_last_error(void);
int dr_device_count(int *out_count);
/* sm count, shared memory per block opt-in, total memory of `device` */
int dr_device_info(int device, int *out_sms, int *out_smem_optin, int64_t *out_total_mem);

/* ---- index lifecycle --------------------------------------------------------------------------
 * Replaces: MMapNodeReader.__init__/get_node (pydiskann/io/diskann_persist.py:209-234),
 *           DiskANNPersist.load_pq_codes / load_pq_codebook (:107-206), and the in-memory
 *           VamanaGraphWithPQ.nodes object graph (vamana_graph.py:8-56).
 * records: N fixed records, float32[D] vector || uint32[R] neighbour ids, exactly the index.dat
 * image written by DiskANNPersist.save_index (:17-24) — rows shorter than R are 0-padded and the
 * padding is honoured as a neighbour, like the reference does.
 * codes u8[N,M] and codebook f32[M,256,D/M] may be NULL (exact search only). */
int dr_index_create_from_records(const void *records, int64_t N, int32_t D, int32_t R,
                                 const uint8_t *codes, const float *codebook, int32_t M,
                                 int64_t medoid, int device, dr_index **out);
/* same, from split host arrays vec f32[N,D], adj u32[N,R] */
int dr_index_create(const float *vec, const uint32_t *adj, const uint8_t *codes, const float *codebook,
                    int64_t N, int32_t D, int32_t R, int32_t M, int64_t medoid, int device, dr_index **out);
/* adopt device arrays (not copied, not freed): for indexes built or generated on the GPU */
int dr_index_create_dev(const float *d_vec, const uint32_t *d_adj, const uint8_t *d_codes,
                        const float *d_codebook, int64_t N, int32_t D, int32_t R, int32_t M, int64_t medoid,
                        int device, dr_index **out);
int dr_index_destroy(dr_index *h);
int dr_index_info(const dr_index *h, int64_t *N, int32_t *D, int32_t *R, int32_t *M, int64_t *medoid, int *device);
/* lazy deletes (VamanaGraphWithPQ.delete_node, vamana_graph.py:116-125; greedy_search_cython skips is_deleted
 * nodes, cython_utils.pyx:109): mask u8[N] on the host, non-zero = deleted; NULL clears it. */
int dr_index_set_deleted(dr_index *h, const uint8_t *mask);
/* change the DEFAULT entry point of the search (a per-call start goes in dr_search_params.start_plus1) */
int dr_index_set_start(dr_index *h, int64_t start);
/* ---- dynamic updates (VamanaGraphWithPQ.insert_node / delete_node, vamana_graph.py:58-125) ----------------------
 * O(rows touched) instead of re-uploading the index.  Only for indexes that own their arrays (not dr_index_create_dev).
 * A row is R slots; a slot holding an id >= N (use 0xFFFFFFFF) is "no neighbour": that is how the live in-memory graph
 * of the reference looks to its searches (they iterate node.neighbors: true degree, set order, no padding —
 * vamana_graph.py:607-640, cython_utils.pyx:108), whereas a row read back from index.dat is 0-padded and the padding IS a
 * neighbour (MMapNodeReader.get_node).  Callers choose by what they write into the unused slots.
 * dr_index_append: n new nodes (ids N .. N+n-1) with vectors vec f32[n,D], codes u8[n,M] (NULL when the index has none),
 *   empty rows; storage grows geometrically.  dr_index_patch_rows: adj u32[n,R] replaces rows[i].  dr_index_patch_vectors:
 *   re-enabling a deleted id with a new vector (:69-74).  dr_index_set_deleted_rows: flags u8[n] for rows[i]. */
int dr_index_append(dr_index *h, const float *vec, const uint8_t *codes, int64_t n);
int dr_index_patch_rows(dr_index *h, const int64_t *rows, int64_t n, const uint32_t *adj);
int dr_index_patch_vectors(dr_index *h, const int64_t *rows, int64_t n, const float *vec, const uint8_t *codes);
int dr_index_set_deleted_rows(dr_index *h, const int64_t *rows, int64_t n, const uint8_t *flags);
/* write the index back as an index.dat image (host buffer of N*4*(D+R) bytes) */
int dr_index_export_records(const dr_index *h, void *records);

/* ---- search -------------------------------------------------------------------------------------
 * Replaces: greedy_search_cython (cython_utils.pyx:72-122) with the ADC callback
 *           (vamana_graph.py:301-329), greedy_search / greedy_search_optimized (vamana_graph.py:607-640,
 *           762-793), beam_search_from_disk (:719-760) and the traversal + exact-distance parts of
 *           SearchEngineCorrect._pq_accelerated_graph_search (search_engine.py:398-506), batched.
 * Q f32[B,D].  lut: NULL (built on the device from the index codebook, bit-identical to
 * DiskANNPQ.compute_distance_table, fast_pq.py:294-318) or f32[B,M,256] supplied by the caller.
 * Outputs (any may be NULL except out_ids): out_ids i32[B,k] (-1 padded), out_dist f32[B,k] (+inf
 * padded), out_hops i32[B] node expansions, out_visited i32[B] distance evaluations,
 * out_list_ids i32[B,L] / out_list_dist f32[B,L] / out_list_len i32[B] the whole final search list
 * with its traversal distances, trace i32[B,trace_cap] ids in the order their distance was computed
 * (W == 1 only), out_status i32[B] DR_ST_* bits. */
int dr_search_batch(dr_index *h, const float *Q, int64_t B, const dr_search_params *p, const float *lut,
                    int32_t *out_ids, float *out_dist, int32_t *out_hops, int32_t *out_visited,
                    int32_t *out_list_ids, float *out_list_dist, int32_t *out_list_len,
                    int32_t *trace, int32_t trace_cap, int32_t *out_status);
/* device-pointer variant; d_lut NULL = build per chunk in an internal buffer.  Enqueues on stream and returns.  The handle's
 * scratch (ADC tables, work counter, overflow tables) is shared by all calls on the handle: a call enqueued on another stream
 * first waits (on the device, cudaStreamWaitEvent) for the previous call's kernels, so calls on one handle never overlap each
 * other; use one handle per concurrent stream to overlap searches. */
int dr_search_batch_dev(dr_index *h, const float *d_Q, int64_t B, const dr_search_params *p, const float *d_lut,
                        int32_t *d_out_ids, float *d_out_dist, int32_t *d_out_hops, int32_t *d_out_visited,
                        int32_t *d_out_list_ids, float *d_out_list_dist, int32_t *d_out_list_len,
                        int32_t *d_trace, int32_t trace_cap, int32_t *d_out_status, void *stream);
/* Variant C with the reference's OWN semantics: beam_search_with_pq (vamana_graph.py:535-605) / beam_search (:690-717), the
 * k-capped beam whose frontier truncation keeps the beam_width WORST entries (:595-596).  dr_search_batch is the search the
 * shims run by default; this entry point returns what the reference's function returns on the same graph, for callers that
 * depend on it (restated in oracle/oracle.c:orc_beam_c and checked against the real reference).  Q f32[B,D] on the host;
 * dist = DR_DIST_PQ (ADC, sequential fp32 sum) or DR_DIST_EXACT (squared L2); start = the call's start_idx;
 * lazily deleted nodes are skipped like :577-584.  out_ids i32[B,k] (-1 padded) and out_dist f32[B,k] sorted by (dist, id);
 * sqrt_out = 1 reports sqrt(d) like :601.  out_dist / out_hops / out_visited may be NULL. */
int dr_beam_search_c(dr_index *h, const float *Q, int64_t B, int32_t k, int32_t beam_width, int32_t dist, int32_t sqrt_out,
                     int64_t start /* -1 = the index's entry point */,
                     int32_t *out_ids, float *out_dist, int32_t *out_hops, int32_t *out_visited);
/* number of kernels this library has launched in this process (bench.py's gpu_launches) */
int64_t dr_launch_count(void);
/* device time (ms) and launches of the search kernel alone, accumulated since the last reset,
 * measured with CUDA events on the launching stream (bench.py's roofline.achieved) */
int dr_search_kernel_timing(dr_index *h, int enable, double *out_ms, int64_t *out_launches);

/* ---- PQ -----------------------------------------------------------------------------------------
 * dr_lut_build replaces DiskANNPQ.compute_distance_table (fast_pq.py:294-318), batched: Q f32[B,D]
 *   -> out f32[B,M,256]; same fp32 operation order as numpy's row reduction (bit-identical). */
int dr_lut_build(dr_index *h, const float *Q, int64_t B, float *out);
int dr_lut_build_dev(dr_index *h, const float *d_Q, int64_t B, float *d_out, void *stream);
/* same without an index: codebook f32[M,256,D/M] on the host (the DiskANNPQ object's own method) */
int dr_pq_lut(const float *codebook, const float *Q, int64_t B, int32_t D, int32_t M, float *out, int device);
/* The throughput search's 8-bit table (DR_LUT_U8 / DR_LUT_U8_TC), exposed for tests: out u8[B, M, 256] in plain
 * (subspace, centroid) order, out_scale / out_offset f32[B]: ADC^2 ~= offset + scale * sum_m out[b][m][code_m]. */
int dr_pq_lut_u8(const float *codebook, const float *Q, int64_t B, int32_t D, int32_t M, int32_t lut_fmt, uint8_t *out,
                 float *out_scale, float *out_offset, int device);
/* dr_pq_train replaces DiskANNPQ.fit (fast_pq.py:197-243; M x KMeans(256)): X f32[N,D] -> codebook
 *   f32[M,256,D/M].  Lloyd iterations on the device; not bit-comparable with sklearn's k-means++
 *    — judged by quantisation error.  out_mse (may be NULL) = mean squared error. */
/* k-means assignment on the tensor cores (tcgen05, TF32) when D / M is a multiple of 8: on by default; 0 forces the exact
 * fp32 CUDA-core assignment (tests compare the two by quantisation error).  The final encode is always exact. */
int dr_pq_train_tensor_cores(int enable);
/* seeding of the codebooks: k-means++ on the device (default; sklearn's init at fast_pq.py:232-240) or, with 0, 256 evenly spaced rows */
int dr_pq_train_kmeanspp(int enable);
int dr_pq_train(const float *X, int64_t N, int32_t D, int32_t M, int32_t iters, uint64_t seed,
                float *out_codebook, double *out_mse, int device);
int dr_pq_train_dev(const float *d_X, int64_t N, int32_t D, int32_t M, int32_t iters, uint64_t seed,
                    float *d_out_codebook, double *out_mse, int device, void *stream);
/* dr_pq_encode replaces DiskANNPQ.encode (fast_pq.py:245-267): nearest centroid, lowest index on ties */
int dr_pq_encode(const float *codebook, const float *X, int64_t N, int32_t D, int32_t M, uint8_t *out_codes, int device);
int dr_pq_encode_dev(const float *d_codebook, const float *d_X, int64_t N, int32_t D, int32_t M, uint8_t *d_out_codes,
                     int device, void *stream);
/* dr_pq_decode replaces DiskANNPQ.decode (fast_pq.py:269-292) */
int dr_pq_decode(const float *codebook, const uint8_t *codes, int64_t N, int32_t D, int32_t M, float *out, int device);
/* dr_adc replaces DiskANNPQ.asymmetric_distance_sq (fast_pq.py:320-328): codes u8[n,M], lut f32[M,256] */
int dr_adc(const uint8_t *codes, const float *lut, int64_t n, int32_t M, float *out, int device);

/* ---- distance helpers ---------------------------------------------------------------------------
 * Replace l2_distance_fast_cython (cython_utils.pyx:18-24), cosine_similarity_cython (:53-70, returns
 * 1 - cos, 0 when a norm is 0) and the dot product, row-wise over n pairs: A,B f32[n,D] -> out f32[n].
 * If nb == 1 the single row B is broadcast against every row of A (query x gathered rows). */
int dr_l2sq_batch(const float *A, const float *B, int64_t n, int64_t nb, int32_t D, float *out, int device);
int dr_dot_batch(const float *A, const float *B, int64_t n, int64_t nb, int32_t D, float *out, int device);
int dr_cosine_batch(const float *A, const float *B, int64_t n, int64_t nb, int32_t D, float *out, int device);
/* symmetric PQ distance, pq_distance_fast_cython (cython_utils.pyx:26-51): c1,c2 u8[n,M] -> f32[n] */
int dr_pq_sdc_batch(const float *codebook, const uint8_t *c1, const uint8_t *c2, int64_t n, int32_t M, int32_t ds,
                    float *out, int device);

/* ---- build --------------------------------------------------------------------------------------
 * dr_medoid replaces compute_approximate_medoid_cython (cython_utils.pyx:210-263) given the sample
 *   ids (ns <= N): argmin_s sum_j ||x_s - x_j||.
 * dr_vamana_build replaces build_vamana_index_cython (cython_utils.pyx:269-492): two passes
 *   (alpha = 1 then alpha), batched greedy search on a snapshot + RobustPrune + reverse-edge merge.
 *   out_adj u32[N,R] 0-padded exactly like DiskANNPersist.save_index, out_deg i32[N] true degrees.
 *   The batched order differs from the reference's sequential insertion, so graphs are compared by
 *   recall, not bit-for-bit. */
int dr_medoid(const float *X, int64_t N, int32_t D, const int32_t *samples, int32_t ns, int64_t *out_medoid, int device);
/* device-pointer form (X and the sample ids already in HBM: indexes generated or built on the GPU) */
int dr_medoid_dev(const float *d_X, int64_t N, int32_t D, const int32_t *d_samples, int32_t ns, int64_t *out_medoid, int device,
                  void *stream);
int dr_vamana_build(const float *X, int64_t N, int32_t D, int32_t R, int32_t L, float alpha, int64_t medoid,
                    uint64_t seed, uint32_t *out_adj, int32_t *out_deg, int device);
int dr_vamana_build_dev(const float *d_X, int64_t N, int32_t D, int32_t R, int32_t L, float alpha, int64_t medoid,
                        uint64_t seed, uint32_t *d_out_adj, int32_t *d_out_deg, int device, void *stream);

/* Status of the last dr_vamana_build[_dev] of this process: how many times a node's incoming reverse-edge run of one batch did not
 * fit the prune kernel's candidate capacity (320) and its tail was dropped (quality only: the dropped edges are reverse edges of
 * hub nodes; 0 on every configuration tested, reported so that it cannot happen silently).  L + R > 320 is refused outright. */
int64_t dr_vamana_build_last_truncated(void);

/* dr_robust_prune replaces robust_prune_cython (cython_utils.pyx:124-167) for one point: p f32[D], cand f32[n,D]
 *   given in ascending id order (n <= 320) -> out_sel i32[<=R] positions into cand in selection order, *out_n. */
int dr_robust_prune(const float *p, const float *cand, int32_t n, int32_t D, float alpha, int32_t R, int32_t *out_sel,
                    int32_t *out_n, int device);

/* ---- multi-GPU merge ------------------------------------------------------------------------------
 * k-way merge of per-shard top-k lists (nothing in the reference): ids i32[G,B,k] (global
 * ids, -1 = empty), dist f32[G,B,k] ascending per shard -> out i32[B,k], f32[B,k].  Device pointers. */
int dr_topk_merge_dev(const int32_t *d_ids, const float *d_dist, int32_t G, int64_t B, int32_t k,
                      int32_t *d_out_ids, float *d_out_dist, int device, void *stream);

/* The same exchange with ONE 64-bit key per entry (f2ord(dist) << 32 | global id, all ones = empty), so that one all-to-all moves
 * ids and distances together.  dr_topk_pack_dev: this rank's [B,k] shard-local lists (ids + id_offset = global) -> send buffer
 * u64[G][Bq][k], Bq = ceil(B / G), block g = the slice of queries rank g reduces.  dr_topk_merge_keys_dev: received u64[G][Bq][k]
 * -> i32[Bq,k], f32[Bq,k]. */
int dr_topk_pack_dev(const int32_t *d_ids, const float *d_dist, int64_t B, int32_t k, int64_t id_offset, int32_t G,
                     uint64_t *d_out_keys, int device, void *stream);
int dr_topk_merge_keys_dev(const uint64_t *d_keys, int32_t G, int64_t Bq, int32_t k, int32_t *d_out_ids, float *d_out_dist,
                           int device, void *stream);
/* The exchange fused into the search kernel: with a peer route set, the throughput kernel's epilogue also writes every query's
 * top-k as packed keys straight into the receive buffer of the rank that reduces that query — peer memory over NVLink / NVSwitch,
 * no collective on the data path — at [rank][row in the owner's slice][k] of the owner's u64[G][Bq][k] buffer.  d_peer_ptrs is a
 * DEVICE array of G pointers (entry g = rank g's receive buffer as mapped into this process: its own allocation for g == rank,
 * dr_ipc_open of the peer's handle otherwise); NULL switches routing off.  Needs rerank = 1, sqrt_out = 0, DR_LUT_U8*. */
int dr_index_set_peer_route(dr_index *h, const uint64_t *d_peer_ptrs, int32_t G, int32_t rank, int64_t B_total, int64_t id_offset);
/* device allocations that can be shared between the ranks of one node (cudaIpc): out_ipc_handle64 receives 64 opaque bytes */
int dr_dev_alloc(int device, int64_t bytes, void **out_ptr, void *out_ipc_handle64);
int dr_dev_free(int device, void *ptr);
int dr_ipc_open(int device, const void *ipc_handle64, void **out_ptr);
int dr_ipc_close(int device, void *ptr);
int dr_dev_memset(int device, void *ptr, int value, int64_t bytes, void *stream);
int dr_dev_upload(int device, void *dst, const void *src_host, int64_t bytes);

#ifdef __cplusplus
}
#endif
#endif /* DISKRAG_B200_H */
