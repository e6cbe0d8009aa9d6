// api.cu — the extern "C" surface of libdiskrag_b200.so (see include/diskrag_b200.h).
#include "common.cuh"

#include <stdarg.h>

#include <vector>

static thread_local std::string g_err;
std::atomic<long long> g_launches{0};

void dr_set_error(const char *fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
}

int dr_scratch(void **ptr, size_t *cur, size_t need) {
    if (*cur >= need) return 0;
    if (*ptr) cudaFree(*ptr);
    *ptr = nullptr;
    *cur = 0;
    DR_CUDA(cudaMalloc(ptr, need));
    *cur = need;
    return 0;
}

// launchers from the other translation units
int launch_adc(const uint8_t *, const float *, int64_t, int, float *, cudaStream_t);
int launch_pq_encode(const float *, const float *, int64_t, int, int, uint8_t *, cudaStream_t);
int launch_pq_train(const float *, int64_t, int, int, int, uint64_t, float *, double *, cudaStream_t);
int launch_pq_decode(const float *, const uint8_t *, int64_t, int, int, float *, cudaStream_t);
int launch_rowdist(const float *, const float *, int64_t, int64_t, int, int, float *, cudaStream_t);
int launch_sdc(const float *, const uint8_t *, const uint8_t *, int64_t, int, int, float *, cudaStream_t);
int launch_topk_merge(const int32_t *, const float *, int, int64_t, int, int32_t *, float *, cudaStream_t);
int launch_topk_pack(const int32_t *, const float *, int64_t, int, int64_t, int, u64 *, cudaStream_t);
int launch_topk_merge_keys(const u64 *, int, int64_t, int, int32_t *, float *, cudaStream_t);
int launch_medoid(const float *, int64_t, int, const int32_t *, int, int, double *, cudaStream_t);
int launch_vamana_build(const float *, int64_t, int, int, int, float, int64_t, uint64_t, uint32_t *, int32_t *, int, cudaStream_t);
int launch_deinterleave(const void *, int64_t, int, int, float *, uint32_t *, cudaStream_t);
int launch_prune_one(const float *, int, int, float, int, uint32_t *, int32_t *, cudaStream_t);
int launch_interleave(const float *, const uint32_t *, int64_t, int, int, void *, cudaStream_t);

extern std::atomic<long long> g_build_truncated;   // build.cu

// small RAII device buffer for the host-pointer entry points
struct DevBuf {
    void *p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
    int alloc(size_t n) { DR_CUDA(cudaMalloc(&p, n ? n : 1)); return 0; }
    template <class T> T *as() { return reinterpret_cast<T *>(p); }
};

static int use_device(int device) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        dr_set_error("no CUDA device available (%s); libdiskrag_b200 has no CPU fallback", cudaGetErrorString(e));
        return 3;
    }
    DR_CHECK(device >= 0 && device < n, "device %d out of range (have %d)", device, n);
    DR_CUDA(cudaSetDevice(device));
    return 0;
}

// records <-> split arrays
__global__ void deinterleave_kernel(const uint32_t *__restrict__ rec, long long N, int D, int R, float *__restrict__ vec,
                                    uint32_t *__restrict__ adj) {
    const int W = D + R;
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= N * W) return;
    long long i = t / W;
    int j = (int)(t - i * W);
    uint32_t v = rec[t];
    if (j < D) vec[(size_t)i * D + j] = __uint_as_float(v);
    else adj[(size_t)i * R + (j - D)] = v;
}
__global__ void interleave_kernel(const float *__restrict__ vec, const uint32_t *__restrict__ adj, long long N, int D, int R,
                                  uint32_t *__restrict__ rec) {
    const int W = D + R;
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= N * W) return;
    long long i = t / W;
    int j = (int)(t - i * W);
    rec[t] = j < D ? __float_as_uint(vec[(size_t)i * D + j]) : adj[(size_t)i * R + (j - D)];
}
int launch_deinterleave(const void *d_rec, int64_t N, int D, int R, float *d_vec, uint32_t *d_adj, cudaStream_t s) {
    long long tot = (long long)N * (D + R);
    if (tot == 0) return 0;
    deinterleave_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, s>>>((const uint32_t *)d_rec, N, D, R, d_vec, d_adj);
    DR_LAUNCHED();
    return 0;
}
int launch_interleave(const float *d_vec, const uint32_t *d_adj, int64_t N, int D, int R, void *d_rec, cudaStream_t s) {
    long long tot = (long long)N * (D + R);
    if (tot == 0) return 0;
    interleave_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, s>>>(d_vec, d_adj, N, D, R, (uint32_t *)d_rec);
    DR_LAUNCHED();
    return 0;
}

// ---- dynamic updates: O(rows touched) -------------------------------------------------------------------------
__global__ void scatter_rows_kernel(const long long *__restrict__ rows, long long n, int words, const uint32_t *__restrict__ src,
                                    uint32_t *__restrict__ dst) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * words) return;
    const long long i = t / words;
    const int j = (int)(t - i * words);
    dst[(size_t)rows[i] * words + j] = src[t];
}
__global__ void scatter_bytes_kernel(const long long *__restrict__ rows, long long n, int width, const uint8_t *__restrict__ src,
                                     uint8_t *__restrict__ dst) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * width) return;
    const long long i = t / width;
    const int j = (int)(t - i * width);
    dst[(size_t)rows[i] * width + j] = src[t];
}
static int scatter_rows(const int64_t *rows_host, int64_t n, int words, const void *src_host, void *dst_dev, int64_t N, bool bytes) {
    if (n == 0) return 0;
    for (int64_t i = 0; i < n; ++i)
        DR_CHECK(rows_host[i] >= 0 && rows_host[i] < N, "dr_index_patch: row %lld out of range (N=%lld)", (long long)rows_host[i], (long long)N);
    DevBuf r, s;
    const size_t payload = (size_t)n * words * (bytes ? 1 : 4);
    if (r.alloc((size_t)n * 8) || s.alloc(payload)) return 1;
    DR_CUDA(cudaMemcpy(r.p, rows_host, (size_t)n * 8, cudaMemcpyHostToDevice));
    DR_CUDA(cudaMemcpy(s.p, src_host, payload, cudaMemcpyHostToDevice));
    const long long tot = (long long)n * words;
    if (bytes) scatter_bytes_kernel<<<(unsigned)((tot + 255) / 256), 256>>>(r.as<long long>(), n, words, s.as<uint8_t>(), (uint8_t *)dst_dev);
    else scatter_rows_kernel<<<(unsigned)((tot + 255) / 256), 256>>>(r.as<long long>(), n, words, s.as<uint32_t>(), (uint32_t *)dst_dev);
    DR_LAUNCHED();
    DR_CUDA(cudaDeviceSynchronize());
    return 0;
}

template <class T>
static int grow_array(T **p, size_t old_elems, size_t new_elems, int fill_byte = -1) {
    if (!*p) return 0;
    T *q = nullptr;
    DR_CUDA(cudaMalloc(&q, new_elems * sizeof(T)));
    DR_CUDA(cudaMemcpy(q, *p, old_elems * sizeof(T), cudaMemcpyDeviceToDevice));
    if (fill_byte >= 0) DR_CUDA(cudaMemset(q + old_elems, fill_byte, (new_elems - old_elems) * sizeof(T)));
    cudaFree(*p);
    *p = q;
    return 0;
}

extern "C" {

int dr_abi_version(void) { return DR_ABI_VERSION; }
const char *dr_last_error(void) { return g_err.c_str(); }
int64_t dr_launch_count(void) { return g_launches.load(); }

int dr_device_count(int *out_count) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) { n = 0; cudaGetLastError(); }
    if (out_count) *out_count = n;
    return 0;
}

int dr_device_info(int device, int *out_sms, int *out_smem_optin, int64_t *out_total_mem) {
    if (use_device(device)) return 3;
    cudaDeviceProp prop;
    DR_CUDA(cudaGetDeviceProperties(&prop, device));
    if (out_sms) *out_sms = prop.multiProcessorCount;
    if (out_smem_optin) *out_smem_optin = (int)prop.sharedMemPerBlockOptin;
    if (out_total_mem) *out_total_mem = (int64_t)prop.totalGlobalMem;
    return 0;
}

static int index_common(dr_index *h, int64_t N, int D, int R, int M, int64_t medoid, int device) {
    DR_CHECK(N > 0 && D > 0 && R > 0, "dr_index: bad shape N=%lld D=%d R=%d", (long long)N, D, R);
    DR_CHECK(N < (1ll << 31), "dr_index: N must be < 2^31");
    DR_CHECK(M == 0 || D % M == 0, "dr_index: D=%d not divisible by M=%d", D, M);
    DR_CHECK(medoid >= 0 && medoid < N, "dr_index: medoid %lld out of range", (long long)medoid);
    h->device = device; h->N = N; h->D = D; h->R = R; h->M = M; h->medoid = medoid;
    cudaDeviceProp prop;
    DR_CUDA(cudaGetDeviceProperties(&prop, device));
    h->sms = prop.multiProcessorCount;
    h->smem_optin = (int)prop.sharedMemPerBlockOptin;
    h->cap = N;
    DR_CUDA(cudaEventCreate(&h->ev0));
    DR_CUDA(cudaEventCreate(&h->ev1));
    DR_CUDA(cudaEventCreateWithFlags(&h->ev_scratch, cudaEventDisableTiming));
    return 0;
}

static int upload_pq(dr_index *h, const uint8_t *codes, const float *codebook) {
    if (h->M > 0 && codes) {
        DR_CUDA(cudaMalloc(&h->d_codes, (size_t)h->N * h->M));
        DR_CUDA(cudaMemcpy(h->d_codes, codes, (size_t)h->N * h->M, cudaMemcpyHostToDevice));
    }
    if (h->M > 0 && codebook) {
        DR_CUDA(cudaMalloc(&h->d_codebook, (size_t)256 * h->D * 4));
        DR_CUDA(cudaMemcpy(h->d_codebook, codebook, (size_t)256 * h->D * 4, cudaMemcpyHostToDevice));
    }
    return 0;
}

int dr_index_create_from_records(const void *records, int64_t N, int32_t D, int32_t R, const uint8_t *codes,
                                 const float *codebook, int32_t M, int64_t medoid, int device, dr_index **out) {
    if (use_device(device)) return 3;
    DR_CHECK(records && out, "dr_index_create_from_records: null argument");
    dr_index *h = new dr_index();
    int rc = index_common(h, N, D, R, M, medoid, device);
    if (!rc) {
        rc = [&]() -> int {
            const size_t bytes = (size_t)N * 4 * (D + R);
            DevBuf rec;
            if (rec.alloc(bytes)) return 1;
            DR_CUDA(cudaMemcpy(rec.p, records, bytes, cudaMemcpyHostToDevice));
            DR_CUDA(cudaMalloc(&h->d_vec, (size_t)N * D * 4));
            DR_CUDA(cudaMalloc(&h->d_adj, (size_t)N * R * 4));
            if (launch_deinterleave(rec.p, N, D, R, h->d_vec, h->d_adj, 0)) return 1;
            DR_CUDA(cudaDeviceSynchronize());
            return upload_pq(h, codes, codebook);
        }();
    }
    if (rc) { dr_index_destroy(h); return rc; }
    *out = h;
    return 0;
}

int dr_index_create(const float *vec, const uint32_t *adj, const uint8_t *codes, const float *codebook, int64_t N, int32_t D,
                    int32_t R, int32_t M, int64_t medoid, int device, dr_index **out) {
    if (use_device(device)) return 3;
    DR_CHECK(vec && adj && out, "dr_index_create: null argument");
    dr_index *h = new dr_index();
    int rc = index_common(h, N, D, R, M, medoid, device);
    if (!rc) {
        rc = [&]() -> int {
            DR_CUDA(cudaMalloc(&h->d_vec, (size_t)N * D * 4));
            DR_CUDA(cudaMalloc(&h->d_adj, (size_t)N * R * 4));
            DR_CUDA(cudaMemcpy(h->d_vec, vec, (size_t)N * D * 4, cudaMemcpyHostToDevice));
            DR_CUDA(cudaMemcpy(h->d_adj, adj, (size_t)N * R * 4, cudaMemcpyHostToDevice));
            return upload_pq(h, codes, codebook);
        }();
    }
    if (rc) { dr_index_destroy(h); return rc; }
    *out = h;
    return 0;
}

int dr_index_create_dev(const float *d_vec, const uint32_t *d_adj, const uint8_t *d_codes, const float *d_codebook, int64_t N,
                        int32_t D, int32_t R, int32_t M, int64_t medoid, int device, dr_index **out) {
    if (use_device(device)) return 3;
    DR_CHECK(d_vec && d_adj && out, "dr_index_create_dev: null argument");
    dr_index *h = new dr_index();
    int rc = index_common(h, N, D, R, M, medoid, device);
    if (rc) { dr_index_destroy(h); return rc; }
    h->owns = false;
    h->d_vec = const_cast<float *>(d_vec);
    h->d_adj = const_cast<uint32_t *>(d_adj);
    h->d_codes = const_cast<uint8_t *>(d_codes);
    h->d_codebook = const_cast<float *>(d_codebook);
    *out = h;
    return 0;
}

int dr_index_destroy(dr_index *h) {
    if (!h) return 0;
    cudaSetDevice(h->device);
    if (h->owns) {
        if (h->d_vec) cudaFree(h->d_vec);
        if (h->d_adj) cudaFree(h->d_adj);
        if (h->d_codes) cudaFree(h->d_codes);
        if (h->d_codebook) cudaFree(h->d_codebook);
    }
    if (h->d_lut) cudaFree(h->d_lut);
    if (h->d_counter) cudaFree(h->d_counter);
    if (h->d_ovf) cudaFree(h->d_ovf);
    if (h->d_io) cudaFree(h->d_io);
    if (h->d_deleted) cudaFree(h->d_deleted);
    if (h->s_in) {
        cudaStreamDestroy(h->s_in); cudaStreamDestroy(h->s_comp); cudaStreamDestroy(h->s_out);
        for (int i = 0; i < DR_PIPE_EVENTS; ++i) { cudaEventDestroy(h->ev_in[i]); cudaEventDestroy(h->ev_done[i]); }
    }
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    if (h->ev_scratch) cudaEventDestroy(h->ev_scratch);
    delete h;
    return 0;
}

int dr_index_info(const dr_index *h, int64_t *N, int32_t *D, int32_t *R, int32_t *M, int64_t *medoid, int *device) {
    DR_CHECK(h, "dr_index_info: null handle");
    if (N) *N = h->N;
    if (D) *D = h->D;
    if (R) *R = h->R;
    if (M) *M = h->M;
    if (medoid) *medoid = h->medoid;
    if (device) *device = h->device;
    return 0;
}

int dr_index_export_records(const dr_index *h, void *records) {
    DR_CHECK(h && records, "dr_index_export_records: null argument");
    DR_LOCK(const_cast<dr_index *>(h));
    DR_CUDA(cudaSetDevice(h->device));
    const size_t bytes = (size_t)h->N * 4 * (h->D + h->R);
    DevBuf rec;
    if (rec.alloc(bytes)) return 1;
    if (launch_interleave(h->d_vec, h->d_adj, h->N, h->D, h->R, rec.p, 0)) return 1;
    DR_CUDA(cudaMemcpy(records, rec.p, bytes, cudaMemcpyDeviceToHost));
    return 0;
}

int dr_search_kernel_timing(dr_index *h, int enable, double *out_ms, int64_t *out_launches) {
    DR_CHECK(h, "dr_search_kernel_timing: null handle");
    DR_LOCK(h);
    if (out_ms) *out_ms = h->timed_ms;
    if (out_launches) *out_launches = h->timed_launches;
    h->timing = enable != 0;
    h->timed_ms = 0.0;
    h->timed_launches = 0;
    return 0;
}

int dr_search_batch_dev(dr_index *h, const float *d_Q, int64_t B, const dr_search_params *p, const float *d_lut,
                        int32_t *d_out_ids, float *d_out_dist, int32_t *d_out_hops, int32_t *d_out_visited,
                        int32_t *d_out_list_ids, float *d_out_list_dist, int32_t *d_out_list_len, int32_t *d_trace,
                        int32_t trace_cap, int32_t *d_out_status, void *stream) {
    DR_CHECK(h && p && d_out_ids && (d_Q || B == 0), "dr_search_batch_dev: null argument");
    DR_LOCK(h);
    DR_CUDA(cudaSetDevice(h->device));
    return launch_search(h, d_Q, B, p, d_lut, d_out_ids, d_out_dist, d_out_hops, d_out_visited, d_out_list_ids,
                         d_out_list_dist, d_out_list_len, d_trace, trace_cap, d_out_status, (cudaStream_t)stream);
}

int dr_search_batch(dr_index *h, const float *Q, int64_t B, const dr_search_params *p, const float *lut, int32_t *out_ids,
                    float *out_dist, int32_t *out_hops, int32_t *out_visited, int32_t *out_list_ids, float *out_list_dist,
                    int32_t *out_list_len, int32_t *trace, int32_t trace_cap, int32_t *out_status) {
    DR_CHECK(h && p && out_ids && (Q || B == 0), "dr_search_batch: null argument");
    DR_CHECK(p->k >= 1 && p->L >= 1 && p->L <= 512, "dr_search_batch: bad k/L");
    DR_LOCK(h);
    DR_CUDA(cudaSetDevice(h->device));
    if (B == 0) return 0;
    // carve one staging allocation: Q | lut? | ids | dist | hops | visited | status | list_ids | list_dist | list_len | trace
    auto al = [](size_t x) { return (x + 255) / 256 * 256; };
    const size_t szQ = al((size_t)B * h->D * 4);
    const size_t szLut = lut ? al((size_t)B * h->M * 1024) : 0;
    const size_t szK = al((size_t)B * p->k * 4), szB = al((size_t)B * 4);
    const size_t szL = al((size_t)B * p->L * 4);
    const size_t szT = trace ? al((size_t)B * trace_cap * 4) : 0;
    size_t total = szQ + szLut + 2 * szK + 4 * szB + 2 * szL + szT;
    if (dr_scratch(&h->d_io, &h->io_bytes, total)) return 1;
    char *base = (char *)h->d_io;
    float *dQ = (float *)base; base += szQ;
    float *dLut = lut ? (float *)base : nullptr; base += szLut;
    int32_t *dIds = (int32_t *)base; base += szK;
    float *dDist = (float *)base; base += szK;
    int32_t *dHops = (int32_t *)base; base += szB;
    int32_t *dVis = (int32_t *)base; base += szB;
    int32_t *dStat = (int32_t *)base; base += szB;
    int32_t *dLlen = (int32_t *)base; base += szB;
    int32_t *dLids = (int32_t *)base; base += szL;
    float *dLdist = (float *)base; base += szL;
    int32_t *dTrace = trace ? (int32_t *)base : nullptr;
    // Pipeline: the batch is cut into pieces; the H2D copy of piece i+1 (copy stream) overlaps the search of piece i
    // (compute stream) and the D2H of piece i-1 (output stream).  With pageable host memory the copies degrade to
    // synchronous staging, which is still correct.
    if (!h->s_in) {
        DR_CUDA(cudaStreamCreateWithFlags(&h->s_in, cudaStreamNonBlocking));
        DR_CUDA(cudaStreamCreateWithFlags(&h->s_comp, cudaStreamNonBlocking));
        DR_CUDA(cudaStreamCreateWithFlags(&h->s_out, cudaStreamNonBlocking));
        for (int i = 0; i < DR_PIPE_EVENTS; ++i) {
            DR_CUDA(cudaEventCreateWithFlags(&h->ev_in[i], cudaEventDisableTiming));
            DR_CUDA(cudaEventCreateWithFlags(&h->ev_done[i], cudaEventDisableTiming));
        }
    }
    // piece sizes: a short ramp (2k, 4k, 8k queries) so that the first search starts after ~0.2 ms of copying instead of a
    // full piece, then equal pieces of about 16k queries (at most DR_PIPE_EVENTS pieces in all)
    int64_t sizes[DR_PIPE_EVENTS];
    int npieces = 0;
    if (!lut && !trace && B >= 8192) {
        int64_t left = B;
        for (int64_t r = 2048; r <= 8192 && left > 4 * r && npieces < 3; r *= 2) { sizes[npieces++] = r; left -= r; }
        int64_t np = (left + 16383) / 16384;
        if (np > DR_PIPE_EVENTS - npieces) np = DR_PIPE_EVENTS - npieces;
        const int64_t each = (left + np - 1) / np;
        while (left > 0) { const int64_t c = left < each ? left : each; sizes[npieces++] = c; left -= c; }
    } else {
        sizes[npieces++] = B;
    }
    if (trace) DR_CUDA(cudaMemsetAsync(dTrace, 0xFF, (size_t)B * trace_cap * 4, h->s_comp));
    int64_t c0 = 0;
    for (int pi = 0; pi < npieces; c0 += sizes[pi], ++pi) {
        const int64_t cb = sizes[pi];
        DR_CUDA(cudaMemcpyAsync(dQ + (size_t)c0 * h->D, Q + (size_t)c0 * h->D, (size_t)cb * h->D * 4, cudaMemcpyHostToDevice, h->s_in));
        if (lut) DR_CUDA(cudaMemcpyAsync(dLut, lut, (size_t)B * h->M * 1024, cudaMemcpyHostToDevice, h->s_in));
        DR_CUDA(cudaEventRecord(h->ev_in[pi], h->s_in));
        DR_CUDA(cudaStreamWaitEvent(h->s_comp, h->ev_in[pi], 0));
        int rc = launch_search(h, dQ + (size_t)c0 * h->D, cb, p, dLut, dIds + (size_t)c0 * p->k, dDist + (size_t)c0 * p->k,
                               dHops + c0, dVis + c0, out_list_ids ? dLids + (size_t)c0 * p->L : nullptr,
                               out_list_dist ? dLdist + (size_t)c0 * p->L : nullptr, dLlen + c0,
                               dTrace ? dTrace + (size_t)c0 * trace_cap : nullptr, trace_cap, dStat + c0, h->s_comp);
        if (rc) { cudaDeviceSynchronize(); return rc; }
        DR_CUDA(cudaEventRecord(h->ev_done[pi], h->s_comp));
        DR_CUDA(cudaStreamWaitEvent(h->s_out, h->ev_done[pi], 0));
        cudaStream_t s = h->s_out;
        DR_CUDA(cudaMemcpyAsync(out_ids + (size_t)c0 * p->k, dIds + (size_t)c0 * p->k, (size_t)cb * p->k * 4, cudaMemcpyDeviceToHost, s));
        if (out_dist) DR_CUDA(cudaMemcpyAsync(out_dist + (size_t)c0 * p->k, dDist + (size_t)c0 * p->k, (size_t)cb * p->k * 4, cudaMemcpyDeviceToHost, s));
        if (out_hops) DR_CUDA(cudaMemcpyAsync(out_hops + c0, dHops + c0, (size_t)cb * 4, cudaMemcpyDeviceToHost, s));
        if (out_visited) DR_CUDA(cudaMemcpyAsync(out_visited + c0, dVis + c0, (size_t)cb * 4, cudaMemcpyDeviceToHost, s));
        if (out_status) DR_CUDA(cudaMemcpyAsync(out_status + c0, dStat + c0, (size_t)cb * 4, cudaMemcpyDeviceToHost, s));
        if (out_list_len) DR_CUDA(cudaMemcpyAsync(out_list_len + c0, dLlen + c0, (size_t)cb * 4, cudaMemcpyDeviceToHost, s));
        if (out_list_ids) DR_CUDA(cudaMemcpyAsync(out_list_ids + (size_t)c0 * p->L, dLids + (size_t)c0 * p->L, (size_t)cb * p->L * 4, cudaMemcpyDeviceToHost, s));
        if (out_list_dist) DR_CUDA(cudaMemcpyAsync(out_list_dist + (size_t)c0 * p->L, dLdist + (size_t)c0 * p->L, (size_t)cb * p->L * 4, cudaMemcpyDeviceToHost, s));
        if (trace) DR_CUDA(cudaMemcpyAsync(trace + (size_t)c0 * trace_cap, dTrace + (size_t)c0 * trace_cap, (size_t)cb * trace_cap * 4, cudaMemcpyDeviceToHost, s));
    }
    DR_CUDA(cudaStreamSynchronize(h->s_out));
    DR_CUDA(cudaStreamSynchronize(h->s_comp));
    return 0;
}

int dr_beam_search_c(dr_index *h, const float *Q, int64_t B, int32_t k, int32_t beam_width, int32_t dist, int32_t sqrt_out,
                     int64_t start, int32_t *out_ids, float *out_dist, int32_t *out_hops, int32_t *out_visited) {
    DR_CHECK(h && out_ids && (Q || B == 0), "dr_beam_search_c: null argument");
    DR_CHECK(k >= 1 && k <= 1024, "dr_beam_search_c: k must be in 1..1024");
    DR_LOCK(h);
    DR_CUDA(cudaSetDevice(h->device));
    if (B == 0) return 0;
    const bool pq = dist == DR_DIST_PQ;
    DR_CHECK(!pq || (h->d_codebook && h->d_codes && h->M > 0), "dr_beam_search_c: index has no PQ codes / codebook");
    const int64_t CH = B < 4096 ? B : 4096;   // queries per launch: bounds the table buffer (M KB per query)
    long long grid; size_t per_cta;
    beam_c_plan(h, CH, &grid, &per_cta);
    DevBuf q, lut, bm, ids, dd, hh, vv;
    if (q.alloc((size_t)CH * h->D * 4) || bm.alloc(per_cta * (size_t)grid) || ids.alloc((size_t)CH * k * 4) ||
        dd.alloc((size_t)CH * k * 4) || hh.alloc((size_t)CH * 4) || vv.alloc((size_t)CH * 4)) return 1;
    if (pq && lut.alloc((size_t)CH * h->M * 1024)) return 1;
    cudaStream_t s = 0;   // one-query-per-call compatibility path: the default stream, no pipeline
    for (int64_t c0 = 0; c0 < B; c0 += CH) {
        const int64_t nb = (B - c0) < CH ? (B - c0) : CH;
        DR_CUDA(cudaMemcpyAsync(q.p, Q + (size_t)c0 * h->D, (size_t)nb * h->D * 4, cudaMemcpyHostToDevice, s));
        if (pq && launch_lut_build(h->d_codebook, q.as<float>(), nb, h->D, h->M, lut.as<float>(), s)) return 1;
        if (launch_beam_c(h, q.as<float>(), nb, k, beam_width, dist, sqrt_out, pq ? lut.as<float>() : nullptr, bm.as<uint32_t>(),
                          ids.as<int32_t>(), dd.as<float>(), hh.as<int32_t>(), vv.as<int32_t>(), s, start)) return 1;
        DR_CUDA(cudaMemcpyAsync(out_ids + (size_t)c0 * k, ids.p, (size_t)nb * k * 4, cudaMemcpyDeviceToHost, s));
        if (out_dist) DR_CUDA(cudaMemcpyAsync(out_dist + (size_t)c0 * k, dd.p, (size_t)nb * k * 4, cudaMemcpyDeviceToHost, s));
        if (out_hops) DR_CUDA(cudaMemcpyAsync(out_hops + c0, hh.p, (size_t)nb * 4, cudaMemcpyDeviceToHost, s));
        if (out_visited) DR_CUDA(cudaMemcpyAsync(out_visited + c0, vv.p, (size_t)nb * 4, cudaMemcpyDeviceToHost, s));
        DR_CUDA(cudaStreamSynchronize(s));
    }
    return 0;
}

int dr_lut_build_dev(dr_index *h, const float *d_Q, int64_t B, float *d_out, void *stream) {
    DR_CHECK(h && h->d_codebook && h->M > 0, "dr_lut_build: index has no codebook");
    DR_LOCK(h);
    DR_CUDA(cudaSetDevice(h->device));
    return launch_lut_build(h->d_codebook, d_Q, B, h->D, h->M, d_out, (cudaStream_t)stream);
}

int dr_lut_build(dr_index *h, const float *Q, int64_t B, float *out) {
    DR_CHECK(h && h->d_codebook && h->M > 0, "dr_lut_build: index has no codebook");
    DR_LOCK(h);
    DR_CUDA(cudaSetDevice(h->device));
    DevBuf q, o;
    if (q.alloc((size_t)B * h->D * 4) || o.alloc((size_t)B * h->M * 1024)) return 1;
    DR_CUDA(cudaMemcpy(q.p, Q, (size_t)B * h->D * 4, cudaMemcpyHostToDevice));
    if (launch_lut_build(h->d_codebook, q.as<float>(), B, h->D, h->M, o.as<float>(), 0)) return 1;
    DR_CUDA(cudaMemcpy(out, o.p, (size_t)B * h->M * 1024, cudaMemcpyDeviceToHost));
    return 0;
}

int dr_pq_lut(const float *codebook, const float *Q, int64_t B, int32_t D, int32_t M, float *out, int device) {
    if (use_device(device)) return 3;
    DR_CHECK(M > 0 && D % M == 0, "dr_pq_lut: D=%d not divisible by M=%d", D, M);
    DevBuf cb, q, o;
    if (cb.alloc((size_t)256 * D * 4) || q.alloc((size_t)B * D * 4) || o.alloc((size_t)B * M * 1024)) return 1;
    DR_CUDA(cudaMemcpy(cb.p, codebook, (size_t)256 * D * 4, cudaMemcpyHostToDevice));
    DR_CUDA(cudaMemcpy(q.p, Q, (size_t)B * D * 4, cudaMemcpyHostToDevice));
    if (launch_lut_build(cb.as<float>(), q.as<float>(), B, D, M, o.as<float>(), 0)) return 1;
    DR_CUDA(cudaMemcpy(out, o.p, (size_t)B * M * 1024, cudaMemcpyDeviceToHost));
    return 0;
}

int dr_pq_lut_u8(const float *codebook, const float *Q, int64_t B, int32_t D, int32_t M, int32_t lut_fmt, uint8_t *out,
                 float *out_scale, float *out_offset, int device) {
    if (use_device(device)) return 3;
    DR_CHECK(M > 0 && D % M == 0, "dr_pq_lut_u8: D=%d not divisible by M=%d", D, M);
    DR_CHECK(lut_fmt == DR_LUT_U8 || lut_fmt == DR_LUT_U8_TC, "dr_pq_lut_u8: lut_fmt must be DR_LUT_U8 or DR_LUT_U8_TC");
    DevBuf cb, q, o, o2, sc, of, mn, rg;
    if (cb.alloc((size_t)256 * D * 4) || q.alloc((size_t)B * D * 4) || o.alloc((size_t)B * M * 256) || o2.alloc((size_t)B * M * 256) ||
        sc.alloc((size_t)B * 4) || of.alloc((size_t)B * 4) || mn.alloc((size_t)B * M * 4) || rg.alloc((size_t)B * 4))
        return 1;
    DR_CUDA(cudaMemcpy(cb.p, codebook, (size_t)256 * D * 4, cudaMemcpyHostToDevice));
    DR_CUDA(cudaMemcpy(q.p, Q, (size_t)B * D * 4, cudaMemcpyHostToDevice));
    const bool words = (M & 3) == 0;
    if (lut_fmt == DR_LUT_U8_TC) {
        int sms = 0;
        DR_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
        if (launch_lut_build_u8_tc(cb.as<float>(), q.as<float>(), B, D, M, o.as<uint8_t>(), sc.as<float>(), of.as<float>(), mn.as<float>(), rg.as<unsigned>(), sms, 0))
            return 1;
    } else if (launch_lut_build_u8(cb.as<float>(), q.as<float>(), B, D, M, o.as<uint8_t>(), sc.as<float>(), of.as<float>(), mn.as<float>(),
                                   rg.as<unsigned>(), words ? 1 : 0, 0)) {
        return 1;
    }
    const uint8_t *plain = o.as<uint8_t>();
    if (words) {
        if (launch_lut_u8_unpermute(o.as<uint8_t>(), B, M, o2.as<uint8_t>(), 0)) return 1;
        plain = o2.as<uint8_t>();
    }
    DR_CUDA(cudaMemcpy(out, plain, (size_t)B * M * 256, cudaMemcpyDeviceToHost));
    DR_CUDA(cudaMemcpy(out_scale, sc.p, (size_t)B * 4, cudaMemcpyDeviceToHost));
    DR_CUDA(cudaMemcpy(out_offset, of.p, (size_t)B * 4, cudaMemcpyDeviceToHost));
    return 0;
}

int dr_pq_train_kmeanspp(int enable) {
    pq_train_set_kmeanspp(enable);
    return 0;
}

int dr_pq_train_tensor_cores(int enable) {
    pq_train_set_tensor_cores(enable);
    return 0;
}

int dr_pq_train_dev(const float *d_X, int64_t N, int32_t D, int32_t M, int32_t iters, uint64_t seed, float *d_out_codebook,
                    double *out_mse, int device, void *stream) {
    if (use_device(device)) return 3;
    return launch_pq_train(d_X, N, D, M, iters, seed, d_out_codebook, out_mse, (cudaStream_t)stream);
}

int dr_pq_train(const float *X, int64_t N, int32_t D, int32_t M, int32_t iters, uint64_t seed, float *out_codebook,
                double *out_mse, int device) {
    if (use_device(device)) return 3;
    DevBuf x, cb;
    if (x.alloc((size_t)N * D * 4) || cb.alloc((size_t)256 * D * 4)) return 1;
    DR_CUDA(cudaMemcpy(x.p, X, (size_t)N * D * 4, cudaMemcpyHostToDevice));
    if (launch_pq_train(x.as<float>(), N, D, M, iters, seed, cb.as<float>(), out_mse, 0)) return 1;
    DR_CUDA(cudaMemcpy(out_codebook, cb.p, (size_t)256 * D * 4, cudaMemcpyDeviceToHost));
    return 0;
}

int dr_pq_encode_dev(const float *d_codebook, const float *d_X, int64_t N, int32_t D, int32_t M, uint8_t *d_out_codes,
                     int device, void *stream) {
    if (use_device(device)) return 3;
    return launch_pq_encode(d_codebook, d_X, N, D, M, d_out_codes, (cudaStream_t)stream);
}

int dr_pq_encode(const float *codebook, const float *X, int64_t N, int32_t D, int32_t M, uint8_t *out_codes, int device) {
    if (use_device(device)) return 3;
    DevBuf x, cb, c;
    if (x.alloc((size_t)N * D * 4) || cb.alloc((size_t)256 * D * 4) || c.alloc((size_t)N * M)) return 1;
    DR_CUDA(cudaMemcpy(x.p, X, (size_t)N * D * 4, cudaMemcpyHostToDevice));
    DR_CUDA(cudaMemcpy(cb.p, codebook, (size_t)256 * D * 4, cudaMemcpyHostToDevice));
    if (launch_pq_encode(cb.as<float>(), x.as<float>(), N, D, M, c.as<uint8_t>(), 0)) return 1;
    DR_CUDA(cudaMemcpy(out_codes, c.p, (size_t)N * M, cudaMemcpyDeviceToHost));
    return 0;
}

int dr_pq_decode(const float *codebook, const uint8_t *codes, int64_t N, int32_t D, int32_t M, float *out, int device) {
    if (use_device(device)) return 3;
    DR_CHECK(M > 0 && D % M == 0, "dr_pq_decode: D=%d not divisible by M=%d", D, M);
    DevBuf cb, c, o;
    if (cb.alloc((size_t)256 * D * 4) || c.alloc((size_t)N * M) || o.alloc((size_t)N * D * 4)) return 1;
    DR_CUDA(cudaMemcpy(cb.p, codebook, (size_t)256 * D * 4, cudaMemcpyHostToDevice));
    DR_CUDA(cudaMemcpy(c.p, codes, (size_t)N * M, cudaMemcpyHostToDevice));
    if (launch_pq_decode(cb.as<float>(), c.as<uint8_t>(), N, D, M, o.as<float>(), 0)) return 1;
    DR_CUDA(cudaMemcpy(out, o.p, (size_t)N * D * 4, cudaMemcpyDeviceToHost));
    return 0;
}

int dr_adc(const uint8_t *codes, const float *lut, int64_t n, int32_t M, float *out, int device) {
    if (use_device(device)) return 3;
    DevBuf c, l, o;
    if (c.alloc((size_t)n * M) || l.alloc((size_t)M * 1024) || o.alloc((size_t)n * 4)) return 1;
    DR_CUDA(cudaMemcpy(c.p, codes, (size_t)n * M, cudaMemcpyHostToDevice));
    DR_CUDA(cudaMemcpy(l.p, lut, (size_t)M * 1024, cudaMemcpyHostToDevice));
    if (launch_adc(c.as<uint8_t>(), l.as<float>(), n, M, o.as<float>(), 0)) return 1;
    DR_CUDA(cudaMemcpy(out, o.p, (size_t)n * 4, cudaMemcpyDeviceToHost));
    return 0;
}

static int rowdist_host(const float *A, const float *B, int64_t n, int64_t nb, int32_t D, int op, float *out, int device) {
    if (use_device(device)) return 3;
    DR_CHECK(nb == n || nb == 1, "row distance: B must have n rows or 1 row (n=%lld nb=%lld)", (long long)n, (long long)nb);
    DevBuf a, b, o;
    if (a.alloc((size_t)n * D * 4) || b.alloc((size_t)nb * D * 4) || o.alloc((size_t)n * 4)) return 1;
    DR_CUDA(cudaMemcpy(a.p, A, (size_t)n * D * 4, cudaMemcpyHostToDevice));
    DR_CUDA(cudaMemcpy(b.p, B, (size_t)nb * D * 4, cudaMemcpyHostToDevice));
    if (launch_rowdist(a.as<float>(), b.as<float>(), n, nb, D, op, o.as<float>(), 0)) return 1;
    DR_CUDA(cudaMemcpy(out, o.p, (size_t)n * 4, cudaMemcpyDeviceToHost));
    return 0;
}
int dr_l2sq_batch(const float *A, const float *B, int64_t n, int64_t nb, int32_t D, float *out, int device) {
    return rowdist_host(A, B, n, nb, D, 0, out, device);
}
int dr_dot_batch(const float *A, const float *B, int64_t n, int64_t nb, int32_t D, float *out, int device) {
    return rowdist_host(A, B, n, nb, D, 1, out, device);
}
int dr_cosine_batch(const float *A, const float *B, int64_t n, int64_t nb, int32_t D, float *out, int device) {
    return rowdist_host(A, B, n, nb, D, 2, out, device);
}

int dr_pq_sdc_batch(const float *codebook, const uint8_t *c1, const uint8_t *c2, int64_t n, int32_t M, int32_t ds, float *out,
                    int device) {
    if (use_device(device)) return 3;
    DevBuf cb, a, b, o;
    if (cb.alloc((size_t)M * 256 * ds * 4) || a.alloc((size_t)n * M) || b.alloc((size_t)n * M) || o.alloc((size_t)n * 4)) return 1;
    DR_CUDA(cudaMemcpy(cb.p, codebook, (size_t)M * 256 * ds * 4, cudaMemcpyHostToDevice));
    DR_CUDA(cudaMemcpy(a.p, c1, (size_t)n * M, cudaMemcpyHostToDevice));
    DR_CUDA(cudaMemcpy(b.p, c2, (size_t)n * M, cudaMemcpyHostToDevice));
    if (launch_sdc(cb.as<float>(), a.as<uint8_t>(), b.as<uint8_t>(), n, M, ds, o.as<float>(), 0)) return 1;
    DR_CUDA(cudaMemcpy(out, o.p, (size_t)n * 4, cudaMemcpyDeviceToHost));
    return 0;
}

int dr_medoid(const float *X, int64_t N, int32_t D, const int32_t *samples, int32_t ns, int64_t *out_medoid, int device) {
    if (use_device(device)) return 3;
    DR_CHECK(ns >= 1 && ns <= N, "dr_medoid: need 1 <= ns <= N");
    DevBuf x, sm, sums;
    if (x.alloc((size_t)N * D * 4) || sm.alloc((size_t)ns * 4) || sums.alloc((size_t)ns * 8)) return 1;
    DR_CUDA(cudaMemcpy(x.p, X, (size_t)N * D * 4, cudaMemcpyHostToDevice));
    DR_CUDA(cudaMemcpy(sm.p, samples, (size_t)ns * 4, cudaMemcpyHostToDevice));
    if (launch_medoid(x.as<float>(), N, D, sm.as<int32_t>(), ns, N <= ns ? 1 : 0, sums.as<double>(), 0)) return 1;
    std::vector<double> hs(ns);
    DR_CUDA(cudaMemcpy(hs.data(), sums.p, (size_t)ns * 8, cudaMemcpyDeviceToHost));
    int best = 0;
    for (int i = 1; i < ns; ++i) if (hs[i] < hs[best]) best = i;
    *out_medoid = samples[best];
    return 0;
}

int dr_medoid_dev(const float *d_X, int64_t N, int32_t D, const int32_t *d_samples, int32_t ns, int64_t *out_medoid, int device,
                  void *stream) {
    if (use_device(device)) return 3;
    DR_CHECK(d_X && d_samples && out_medoid && ns >= 1 && ns <= N, "dr_medoid_dev: need device arrays and 1 <= ns <= N");
    DevBuf sums;
    if (sums.alloc((size_t)ns * 8)) return 1;
    cudaStream_t s = (cudaStream_t)stream;
    if (launch_medoid(d_X, N, D, d_samples, ns, N <= ns ? 1 : 0, sums.as<double>(), s)) return 1;
    std::vector<double> hs(ns);
    std::vector<int32_t> smp(ns);
    DR_CUDA(cudaMemcpyAsync(hs.data(), sums.p, (size_t)ns * 8, cudaMemcpyDeviceToHost, s));
    DR_CUDA(cudaMemcpyAsync(smp.data(), d_samples, (size_t)ns * 4, cudaMemcpyDeviceToHost, s));
    DR_CUDA(cudaStreamSynchronize(s));
    int best = 0;
    for (int i = 1; i < ns; ++i) if (hs[i] < hs[best]) best = i;
    *out_medoid = smp[best];
    return 0;
}

int dr_vamana_build_dev(const float *d_X, int64_t N, int32_t D, int32_t R, int32_t L, float alpha, int64_t medoid,
                        uint64_t seed, uint32_t *d_out_adj, int32_t *d_out_deg, int device, void *stream) {
    if (use_device(device)) return 3;
    return launch_vamana_build(d_X, N, D, R, L, alpha, medoid, seed, d_out_adj, d_out_deg, device, (cudaStream_t)stream);
}

int dr_vamana_build(const float *X, int64_t N, int32_t D, int32_t R, int32_t L, float alpha, int64_t medoid, uint64_t seed,
                    uint32_t *out_adj, int32_t *out_deg, int device) {
    if (use_device(device)) return 3;
    DevBuf x, adj, deg;
    if (x.alloc((size_t)N * D * 4) || adj.alloc((size_t)N * R * 4) || deg.alloc((size_t)N * 4)) return 1;
    DR_CUDA(cudaMemcpy(x.p, X, (size_t)N * D * 4, cudaMemcpyHostToDevice));
    if (launch_vamana_build(x.as<float>(), N, D, R, L, alpha, medoid, seed, adj.as<uint32_t>(), deg.as<int32_t>(), device, 0)) return 1;
    DR_CUDA(cudaDeviceSynchronize());
    DR_CUDA(cudaMemcpy(out_adj, adj.p, (size_t)N * R * 4, cudaMemcpyDeviceToHost));
    if (out_deg) DR_CUDA(cudaMemcpy(out_deg, deg.p, (size_t)N * 4, cudaMemcpyDeviceToHost));
    return 0;
}

int64_t dr_vamana_build_last_truncated(void) { return g_build_truncated.load(); }

int dr_robust_prune(const float *p, const float *cand, int32_t n, int32_t D, float alpha, int32_t R, int32_t *out_sel,
                    int32_t *out_n, int device) {
    if (use_device(device)) return 3;
    DR_CHECK(p && out_sel && out_n && (cand || n == 0), "dr_robust_prune: null argument");
    const int stride = n > R ? n : R;
    DevBuf x, row, deg;
    if (x.alloc((size_t)(n + 1) * D * 4) || row.alloc((size_t)stride * 4) || deg.alloc(4)) return 1;
    if (n) DR_CUDA(cudaMemcpy(x.p, cand, (size_t)n * D * 4, cudaMemcpyHostToDevice));
    DR_CUDA(cudaMemcpy((char *)x.p + (size_t)n * D * 4, p, (size_t)D * 4, cudaMemcpyHostToDevice));
    std::vector<uint32_t> ids(stride, 0);
    for (int i = 0; i < n; ++i) ids[i] = (uint32_t)i;
    DR_CUDA(cudaMemcpy(row.p, ids.data(), (size_t)stride * 4, cudaMemcpyHostToDevice));
    DR_CUDA(cudaMemcpy(deg.p, &n, 4, cudaMemcpyHostToDevice));
    if (launch_prune_one(x.as<float>(), n, D, alpha, R, row.as<uint32_t>(), deg.as<int32_t>(), 0)) return 1;
    int cnt = 0;
    DR_CUDA(cudaMemcpy(&cnt, deg.p, 4, cudaMemcpyDeviceToHost));
    DR_CUDA(cudaMemcpy(ids.data(), row.p, (size_t)stride * 4, cudaMemcpyDeviceToHost));
    for (int i = 0; i < cnt; ++i) out_sel[i] = (int32_t)ids[i];
    *out_n = cnt;
    return 0;
}

int dr_index_set_deleted(dr_index *h, const uint8_t *mask) {
    DR_CHECK(h, "dr_index_set_deleted: null handle");
    DR_LOCK(h);
    DR_CUDA(cudaSetDevice(h->device));
    if (!mask) { if (h->d_deleted) cudaFree(h->d_deleted); h->d_deleted = nullptr; return 0; }
    if (!h->d_deleted) DR_CUDA(cudaMalloc(&h->d_deleted, (size_t)(h->cap > h->N ? h->cap : h->N)));
    DR_CUDA(cudaMemcpy(h->d_deleted, mask, (size_t)h->N, cudaMemcpyHostToDevice));
    return 0;
}

int dr_index_append(dr_index *h, const float *vec, const uint8_t *codes, int64_t n) {
    DR_CHECK(h && (vec || n == 0) && n >= 0, "dr_index_append: bad argument");
    DR_LOCK(h);
    DR_CUDA(cudaSetDevice(h->device));
    DR_CHECK(h->owns, "dr_index_append: the index adopted caller-owned device arrays (dr_index_create_dev) and cannot grow");
    DR_CHECK(!(h->d_codes && !codes && n > 0), "dr_index_append: the index has PQ codes; codes u8[n,M] are required");
    DR_CHECK(h->N + n < (1ll << 31), "dr_index_append: N must stay < 2^31");
    if (n == 0) return 0;
    DR_CUDA(cudaDeviceSynchronize());                    // no search may still be reading the arrays that are about to move
    if (h->N + n > h->cap) {
        int64_t ncap = h->cap + h->cap / 2 + 1024;
        if (ncap < h->N + n) ncap = h->N + n;
        if (grow_array(&h->d_vec, (size_t)h->N * h->D, (size_t)ncap * h->D)) return 1;
        if (grow_array(&h->d_adj, (size_t)h->N * h->R, (size_t)ncap * h->R, 0xFF)) return 1;      // new rows: no neighbours
        if (grow_array(&h->d_codes, (size_t)h->N * h->M, (size_t)ncap * h->M)) return 1;
        if (grow_array(&h->d_deleted, (size_t)h->N, (size_t)ncap, 0)) return 1;
        h->cap = ncap;
    } else {
        DR_CUDA(cudaMemset(h->d_adj + (size_t)h->N * h->R, 0xFF, (size_t)n * h->R * 4));
        if (h->d_deleted) DR_CUDA(cudaMemset(h->d_deleted + h->N, 0, (size_t)n));
    }
    DR_CUDA(cudaMemcpy(h->d_vec + (size_t)h->N * h->D, vec, (size_t)n * h->D * 4, cudaMemcpyHostToDevice));
    if (h->d_codes) DR_CUDA(cudaMemcpy(h->d_codes + (size_t)h->N * h->M, codes, (size_t)n * h->M, cudaMemcpyHostToDevice));
    h->N += n;
    return 0;
}

int dr_index_patch_rows(dr_index *h, const int64_t *rows, int64_t n, const uint32_t *adj) {
    DR_CHECK(h && (n == 0 || (rows && adj)), "dr_index_patch_rows: null argument");
    DR_LOCK(h);
    DR_CUDA(cudaSetDevice(h->device));
    DR_CHECK(h->owns, "dr_index_patch_rows: the index adopted caller-owned device arrays (dr_index_create_dev)");
    DR_CUDA(cudaDeviceSynchronize());
    return scatter_rows(rows, n, h->R, adj, h->d_adj, h->N, false);
}

int dr_index_patch_vectors(dr_index *h, const int64_t *rows, int64_t n, const float *vec, const uint8_t *codes) {
    DR_CHECK(h && (n == 0 || (rows && vec)), "dr_index_patch_vectors: null argument");
    DR_LOCK(h);
    DR_CUDA(cudaSetDevice(h->device));
    DR_CHECK(h->owns, "dr_index_patch_vectors: the index adopted caller-owned device arrays (dr_index_create_dev)");
    DR_CUDA(cudaDeviceSynchronize());
    if (scatter_rows(rows, n, h->D, vec, h->d_vec, h->N, false)) return 1;
    if (h->d_codes && codes) return scatter_rows(rows, n, h->M, codes, h->d_codes, h->N, true);
    return 0;
}

int dr_index_set_deleted_rows(dr_index *h, const int64_t *rows, int64_t n, const uint8_t *flags) {
    DR_CHECK(h && (n == 0 || (rows && flags)), "dr_index_set_deleted_rows: null argument");
    DR_LOCK(h);
    DR_CUDA(cudaSetDevice(h->device));
    DR_CUDA(cudaDeviceSynchronize());
    if (!h->d_deleted) {
        DR_CUDA(cudaMalloc(&h->d_deleted, (size_t)(h->cap > h->N ? h->cap : h->N)));
        DR_CUDA(cudaMemset(h->d_deleted, 0, (size_t)(h->cap > h->N ? h->cap : h->N)));
    }
    return scatter_rows(rows, n, 1, flags, h->d_deleted, h->N, true);
}

int dr_index_set_start(dr_index *h, int64_t start) {
    DR_CHECK(h && start >= 0 && start < h->N, "dr_index_set_start: out of range");
    DR_LOCK(h);
    h->medoid = start;
    return 0;
}

int dr_topk_merge_dev(const int32_t *d_ids, const float *d_dist, int32_t G, int64_t B, int32_t k, int32_t *d_out_ids,
                      float *d_out_dist, int device, void *stream) {
    if (use_device(device)) return 3;
    return launch_topk_merge(d_ids, d_dist, G, B, k, d_out_ids, d_out_dist, (cudaStream_t)stream);
}

int dr_topk_pack_dev(const int32_t *d_ids, const float *d_dist, int64_t B, int32_t k, int64_t id_offset, int32_t G, uint64_t *d_out_keys,
                     int device, void *stream) {
    if (use_device(device)) return 3;
    return launch_topk_pack(d_ids, d_dist, B, k, id_offset, G, (u64 *)d_out_keys, (cudaStream_t)stream);
}

int dr_topk_merge_keys_dev(const uint64_t *d_keys, int32_t G, int64_t Bq, int32_t k, int32_t *d_out_ids, float *d_out_dist, int device,
                           void *stream) {
    if (use_device(device)) return 3;
    return launch_topk_merge_keys((const u64 *)d_keys, G, Bq, k, d_out_ids, d_out_dist, (cudaStream_t)stream);
}

int dr_index_set_peer_route(dr_index *h, const uint64_t *d_peer_ptrs, int32_t G, int32_t rank, int64_t B_total, int64_t id_offset) {
    DR_CHECK(h, "dr_index_set_peer_route: null handle");
    DR_LOCK(h);
    if (!d_peer_ptrs) { h->d_peer_recv = nullptr; h->peer_G = 0; return 0; }
    DR_CHECK(G >= 1 && rank >= 0 && rank < G && B_total >= 1 && id_offset >= 0, "dr_index_set_peer_route: bad G / rank / batch / offset");
    h->d_peer_recv = reinterpret_cast<u64 *const *>(d_peer_ptrs);
    h->peer_G = G; h->peer_rank = rank; h->peer_B = B_total; h->peer_id_offset = id_offset;
    return 0;
}

// plain device allocations with a process-portable handle: symmetric receive buffers for the peer-routed exchange
int dr_dev_alloc(int device, int64_t bytes, void **out_ptr, void *out_ipc_handle64) {
    if (use_device(device)) return 3;
    DR_CHECK(out_ptr && bytes > 0, "dr_dev_alloc: bad argument");
    DR_CUDA(cudaMalloc(out_ptr, (size_t)bytes));
    if (out_ipc_handle64) {
        cudaIpcMemHandle_t hd;
        DR_CUDA(cudaIpcGetMemHandle(&hd, *out_ptr));
        static_assert(sizeof(hd) == 64, "cudaIpcMemHandle_t is 64 bytes");
        memcpy(out_ipc_handle64, &hd, 64);
    }
    return 0;
}
int dr_dev_free(int device, void *ptr) {
    if (use_device(device)) return 3;
    if (ptr) DR_CUDA(cudaFree(ptr));
    return 0;
}
int dr_ipc_open(int device, const void *ipc_handle64, void **out_ptr) {
    if (use_device(device)) return 3;
    DR_CHECK(ipc_handle64 && out_ptr, "dr_ipc_open: null argument");
    cudaIpcMemHandle_t hd;
    memcpy(&hd, ipc_handle64, 64);
    DR_CUDA(cudaIpcOpenMemHandle(out_ptr, hd, cudaIpcMemLazyEnablePeerAccess));
    return 0;
}
int dr_ipc_close(int device, void *ptr) {
    if (use_device(device)) return 3;
    if (ptr) DR_CUDA(cudaIpcCloseMemHandle(ptr));
    return 0;
}
int dr_dev_memset(int device, void *ptr, int value, int64_t bytes, void *stream) {
    if (use_device(device)) return 3;
    DR_CUDA(cudaMemsetAsync(ptr, value, (size_t)bytes, (cudaStream_t)stream));
    return 0;
}
int dr_dev_upload(int device, void *dst, const void *src_host, int64_t bytes) {
    if (use_device(device)) return 3;
    DR_CUDA(cudaMemcpy(dst, src_host, (size_t)bytes, cudaMemcpyHostToDevice));
    return 0;
}

}  // extern "C"
