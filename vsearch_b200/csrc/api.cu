// api.cu -- the extern "C" boundary declared in include/vsearch_b200.h.
#include <stdarg.h>
#include <stdlib.h>

#include <vector>

#include "index.cuh"

namespace vs {

static thread_local char g_err[1024] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

// scan.cu / merge.cu
size_t scan_smem_bytes(int vpad, int cap);
int scan_cap_for_k(int k);
int scan_kout(const vs_index *idx, int k);
int launch_prep_query(const void *d_q, int q_dtype, int64_t B, int64_t ldq, int64_t n_cols, int vpad, int round_mode,
                      float *d_out, cudaStream_t st);
int launch_scan(const vs_index *idx, const float *d_qprep, int vpad, int64_t B, int k, int score_round,
                uint64_t *d_cand, float *d_scores_out, cudaStream_t st);
size_t inverted_workspace_bytes(const vs_index *idx, int64_t Bc, int group);
int inverted_extract(vs_index *idx, const float *d_qprep, int vpad, int64_t Bc, void *d_ws, uint32_t *max_nnz,
                     double *mean_postings, uint64_t *max_postings, cudaStream_t st);
bool inverted_usable(uint32_t max_nnz, uint64_t max_postings);
int launch_inverted(vs_index *idx, int64_t Bc, int k, int score_round, int group, uint32_t max_nnz, void *d_ws,
                    uint64_t *d_cand, cudaEvent_t ev0, cudaEvent_t ev1, cudaStream_t st);
size_t dense_workspace_bytes(const vs_index *idx, int64_t B, int k);
int search_dense(vs_index *idx, const void *d_q, int q_dtype, int64_t B, int64_t ldq, int k, int score_round,
                 int64_t id_offset, int64_t *d_ids, float *d_scores, uint64_t *d_keys, void *d_ws, cudaStream_t st);
int launch_merge(const uint64_t *d_in, int64_t P, int64_t stride_p, int64_t stride_b, int64_t B, int k_in, int k_out,
                 int64_t id_offset, int64_t *d_ids, float *d_scores, uint64_t *d_keys, cudaStream_t st);

static size_t dtype_size(int dt) {
    switch (dt) {
        case VS_F32: case VS_I32: case VS_U32: return 4;
        case VS_F16: case VS_BF16: case VS_U16: return 2;
        case VS_I64: return 8;
        default: return 0;
    }
}

static bool is_device_ptr(const void *p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

// RAII device staging copy of a host buffer
struct Staged {
    const void *ptr = nullptr;
    void *owned = nullptr;
    ~Staged() { if (owned) cudaFree(owned); }
    int init(const void *hd, size_t bytes, cudaStream_t st) {
        if (hd == nullptr || bytes == 0 || is_device_ptr(hd)) { ptr = hd; return VS_OK; }
        VS_CUDA(cudaMalloc(&owned, bytes));
        VS_CUDA(cudaMemcpyAsync(owned, hd, bytes, cudaMemcpyHostToDevice, st));
        ptr = owned;
        return VS_OK;
    }
};

static inline int vpad_for(int64_t n_cols) { return (int)(((n_cols + 1) + 3) / 4 * 4); }

// workspace layout for one query chunk of Bc queries
struct Workspace {
    float *qprep;      // [Bc, vpad]
    uint64_t *cand;    // [Bc, n_ctas, k]
    void *inv;         // K3: query lists + accumulators
    size_t bytes;
};
constexpr int kInvGroup = 1;   // queries scored concurrently by the inverted-list path (accumulator rows)
static size_t align256(size_t x) { return (x + 255) / 256 * 256; }
static Workspace carve(const vs_index *idx, void *base, int64_t Bc, int k) {
    Workspace w;
    size_t q_bytes = align256((size_t)Bc * vpad_for(idx->n_cols) * 4);
    size_t c_bytes = align256((size_t)Bc * idx->n_ctas * (size_t)scan_kout(idx, k) * 8);   // scan lists are the longer ones
    w.qprep = (float *)base;
    w.cand = (uint64_t *)((uint8_t *)base + q_bytes);
    w.inv = (uint8_t *)base + q_bytes + c_bytes;
    w.bytes = q_bytes + c_bytes + align256(inverted_workspace_bytes(idx, Bc, kInvGroup));
    return w;
}
constexpr int64_t kQueryChunk = 1024;
// queries per launch: at most kQueryChunk, fewer when the per-CTA candidate lists of a chunk would pass 512 MB (large k)
static int64_t query_chunk(const vs_index *idx, int64_t B, int k) {
    const int64_t per_query = (int64_t)idx->n_ctas * scan_kout(idx, k) * 8;
    int64_t c = (512ll << 20) / (per_query > 0 ? per_query : 1);
    c = c < 16 ? 16 : (c > kQueryChunk ? kQueryChunk : c);
    return B < c ? B : c;
}
// auto-mode cost model; refined from measurements (profiles/)
constexpr double kScanBytesPerSecPair = 4.5e12;   // binary / 16-bit values (L1 data-pipe bound)
constexpr double kScanBytesPerSecF32 = 6.2e12;    // fp32 values (HBM bound)
constexpr double kInvPostingsPerSec = 3.5e11;     // shared-memory atomics, all SMs
constexpr double kInvSecPerRow = 1.8e-12;         // zero + select of the block accumulators
constexpr double kInvFixedSec = 5.0e-6;


// ---- score given rows (rerank stage) ---------------------------------------------------------
// One warp per (query, candidate): the candidate's run of 16-byte chunks in the WS stream (row_chunk[]), lanes stride
// the chunks, q[col] gathered from the prepared query in global memory (L2-resident), fp32 accumulate.
__global__ void __launch_bounds__(256) score_rows_kernel(const WsView idx, const float *qprep, int vpad, const int64_t *ids,
                                                         int64_t n_pairs, int k, int score_round, float *out) {
    const int lane = threadIdx.x & 31;
    const int64_t pair = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (pair >= n_pairs) return;
    const int64_t id = ids[pair];
    if (id < 0 || id >= idx.n_rows) {
        if (lane == 0) out[pair] = -INFINITY;
        return;
    }
    const float *q = qprep + (pair / k) * (int64_t)vpad;
    const uint32_t c0 = idx.row_chunk[id], c1 = idx.row_chunk[id + 1];
    float s = 0.f;
    for (uint32_t c = c0 + lane; c < c1; c += 32) {
        const uint64_t pc = ws_phys_chunk(c, idx.cpl_shift);   // c runs over the row's LOGICAL chunks
        const uint4 u = idx.cols[pc];
        const uint32_t wv[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const uint32_t col = ((e & 1) ? (wv[e >> 1] >> 16) : (wv[e >> 1] & 0xffffu)) & 0x7fffu;   // bit 15 = row-end flag
            float v = 1.f;   // binary index; padding entries point at the query's zero slot
            if (idx.kind == 1) {
                const uint64_t at = pc * 8ull + e;
                if (idx.store_dtype == VS_F32) v = ((const float *)idx.vals)[at];
                else if (idx.store_dtype == VS_F16) v = __half2float(((const __half *)idx.vals)[at]);
                else v = __bfloat162float(((const __nv_bfloat16 *)idx.vals)[at]);
            }
            s += q[col] * v;
        }
    }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
    if (lane == 0) out[pair] = round_score(s, score_round);
}

}  // namespace vs

using namespace vs;

extern "C" {

const char *vs_last_error(void) { return vs::g_err; }
int vs_abi_version(void) { return VS_ABI_VERSION; }

int vs_index_create_csr(int device, int64_t n_rows, int64_t n_cols, int64_t nnz, const void *hd_crow, int crow_dtype,
                        const void *hd_col, int col_dtype, const void *hd_val, int val_dtype, int store_dtype,
                        void *stream, vs_index **out) {
    VS_REQUIRE(out != nullptr, VS_ERR_INVALID, "out is NULL");
    *out = nullptr;
    VS_REQUIRE(n_rows >= 0 && n_cols >= 1 && nnz >= 0, VS_ERR_INVALID, "bad shape");
    VS_REQUIRE(n_cols <= 32767, VS_ERR_UNSUPPORTED,
               "n_cols=%lld does not fit the uint16 column format (15 bits + the row-end flag)", (long long)n_cols);
    VS_REQUIRE(n_rows < 0xffffffffll, VS_ERR_UNSUPPORTED, "n_rows must be < 2^32 - 1 per shard");
    VS_REQUIRE(crow_dtype == VS_I32 || crow_dtype == VS_I64, VS_ERR_INVALID, "crow dtype must be int32/int64");
    VS_REQUIRE(col_dtype == VS_I32 || col_dtype == VS_I64, VS_ERR_INVALID, "col dtype must be int32/int64");
    const bool binary = (hd_val == nullptr || val_dtype == VS_NONE);
    if (!binary) {
        VS_REQUIRE(val_dtype == VS_F32 || val_dtype == VS_F16 || val_dtype == VS_BF16, VS_ERR_INVALID, "bad value dtype");
        VS_REQUIRE(store_dtype == VS_F32 || store_dtype == VS_F16 || store_dtype == VS_BF16, VS_ERR_INVALID, "bad store dtype");
    }
    VS_REQUIRE(hd_crow != nullptr && (nnz == 0 || hd_col != nullptr), VS_ERR_INVALID, "NULL CSR arrays");
    VS_CUDA(cudaSetDevice(device));
    cudaStream_t st = (cudaStream_t)stream;

    vs_index *idx = new (std::nothrow) vs_index();
    VS_REQUIRE(idx != nullptr, VS_ERR_NOMEM, "out of host memory");
    idx->device = device;
    idx->kind = binary ? 2 : 1;
    idx->store_dtype = binary ? VS_NONE : store_dtype;
    idx->n_rows = n_rows; idx->n_cols = n_cols; idx->nnz = nnz;
    if (const char *e = getenv("VSEARCH_B200_BANK_AWARE")) idx->bank_aware = (e[0] != '0');

    int rc;
    {
        Staged s_crow, s_col, s_val;
        rc = s_crow.init(hd_crow, (size_t)(n_rows + 1) * dtype_size(crow_dtype), st);
        if (rc == VS_OK) rc = s_col.init(hd_col, (size_t)nnz * dtype_size(col_dtype), st);
        if (rc == VS_OK && !binary) rc = s_val.init(hd_val, (size_t)nnz * dtype_size(val_dtype), st);
        if (rc == VS_OK) rc = build_ws_index(idx, s_crow.ptr, crow_dtype, s_col.ptr, col_dtype, binary ? nullptr : s_val.ptr,
                                             binary ? VS_NONE : val_dtype, st);
        cudaStreamSynchronize(st);
    }
    if (rc == VS_OK) {
        for (int i = 0; i < VS_TIMER_SLOTS && rc == VS_OK; ++i)
            if (cudaEventCreate(&idx->ev0[i]) != cudaSuccess || cudaEventCreate(&idx->ev1[i]) != cudaSuccess) {
                set_error("cudaEventCreate failed");
                rc = VS_ERR_CUDA;
            }
    }
    if (rc != VS_OK) { vs_index_destroy(idx); return rc; }
    *out = idx;
    return VS_OK;
}

int vs_index_create_dense(int device, int64_t n_rows, int64_t dim, const void *hd_x, int x_dtype, int64_t ld,
                          int store_dtype, void *stream, vs_index **out) {
    VS_REQUIRE(out != nullptr, VS_ERR_INVALID, "out is NULL");
    *out = nullptr;
    VS_REQUIRE(n_rows >= 0 && dim >= 1 && ld >= dim && hd_x != nullptr, VS_ERR_INVALID, "bad dense shape");
    VS_REQUIRE(x_dtype == VS_F32 || x_dtype == VS_F16 || x_dtype == VS_BF16, VS_ERR_INVALID, "bad dense dtype");
    VS_REQUIRE(store_dtype == VS_F16 || store_dtype == VS_BF16, VS_ERR_UNSUPPORTED,
               "the dense index is stored as bf16 or fp16 (tcgen05 kind::f16); fp32 storage is not built");
    VS_REQUIRE(n_rows < 0x7fffff00ll, VS_ERR_UNSUPPORTED, "n_rows must be < 2^31 per shard");
    VS_CUDA(cudaSetDevice(device));
    cudaStream_t st = (cudaStream_t)stream;
    vs_index *idx = new (std::nothrow) vs_index();
    VS_REQUIRE(idx != nullptr, VS_ERR_NOMEM, "out of host memory");
    idx->device = device; idx->kind = 0; idx->store_dtype = store_dtype;
    idx->n_rows = n_rows; idx->n_cols = dim; idx->dim = dim; idx->nnz = n_rows * dim;
    int rc;
    {
        Staged sx;
        rc = sx.init(hd_x, (size_t)n_rows * ld * dtype_size(x_dtype), st);
        if (rc == VS_OK) rc = build_dense_index(idx, sx.ptr, x_dtype, ld, st);
        cudaStreamSynchronize(st);
    }
    for (int i = 0; i < VS_TIMER_SLOTS && rc == VS_OK; ++i)
        if (cudaEventCreate(&idx->ev0[i]) != cudaSuccess || cudaEventCreate(&idx->ev1[i]) != cudaSuccess) {
            set_error("cudaEventCreate failed");
            rc = VS_ERR_CUDA;
        }
    if (rc != VS_OK) { vs_index_destroy(idx); return rc; }
    *out = idx;
    return VS_OK;
}

int vs_index_destroy(vs_index *idx) {
    if (!idx) return VS_OK;
    cudaSetDevice(idx->device);
    cudaFree(idx->cols); cudaFree(idx->vals); cudaFree(idx->tails);
    cudaFree(idx->part_win_begin); cudaFree(idx->part_row_begin); cudaFree(idx->row_chunk);
    cudaFree(idx->dense);
    cudaFree(idx->post_ptr); cudaFree(idx->blk_ptr); cudaFree(idx->blk_base); cudaFree(idx->post_row); cudaFree(idx->post_val);
    for (int i = 0; i < VS_TIMER_SLOTS; ++i) {
        if (idx->ev0[i]) cudaEventDestroy(idx->ev0[i]);
        if (idx->ev1[i]) cudaEventDestroy(idx->ev1[i]);
    }
    delete idx;
    return VS_OK;
}

int vs_index_info(const vs_index *idx, int64_t *n_rows, int64_t *n_cols, int64_t *nnz, int *kind, int *store_dtype,
                  int64_t *device_bytes, int64_t *stream_bytes) {
    VS_REQUIRE(idx != nullptr, VS_ERR_INVALID, "index is NULL");
    if (n_rows) *n_rows = idx->n_rows;
    if (n_cols) *n_cols = idx->n_cols;
    if (nnz) *nnz = idx->nnz;
    if (kind) *kind = idx->kind;
    if (store_dtype) *store_dtype = idx->store_dtype;
    if (device_bytes) *device_bytes = idx->device_bytes;
    if (stream_bytes) *stream_bytes = idx->stream_bytes;
    return VS_OK;
}

int vs_index_export_csr(const vs_index *idx, int64_t *d_crow, int64_t *d_col, float *d_val, void *stream) {
    VS_REQUIRE(idx != nullptr && idx->kind != 0, VS_ERR_INVALID, "export needs a sparse/binary index");
    VS_CUDA(cudaSetDevice(idx->device));
    return export_ws_csr(idx, d_crow, d_col, d_val, (cudaStream_t)stream);
}

size_t vs_search_workspace_bytes(const vs_index *idx, int64_t B, int k) {
    if (!idx || B <= 0 || k <= 0) return 256;
    if (idx->kind == 0) return dense_workspace_bytes(idx, B, k) + 256;
    return carve(idx, nullptr, query_chunk(idx, B, k), k).bytes + 256;
}

static int search_impl(const vs_index *cidx, const void *hd_q, int q_dtype, int64_t B, int64_t ldq, int k, int mode,
                       int score_round, int64_t id_offset, int64_t *d_ids, float *d_scores, uint64_t *d_keys,
                       float *d_scores_full, void *d_workspace, size_t workspace_bytes, void *stream) {
    vs_index *idx = const_cast<vs_index *>(cidx);
    VS_REQUIRE(idx != nullptr, VS_ERR_INVALID, "index is NULL");
    VS_REQUIRE(B >= 0 && ldq >= idx->n_cols, VS_ERR_INVALID, "query leading dimension %lld < n_cols %lld", (long long)ldq,
               (long long)idx->n_cols);
    VS_REQUIRE(q_dtype == VS_F32 || q_dtype == VS_F16 || q_dtype == VS_BF16, VS_ERR_INVALID, "bad query dtype");
    VS_REQUIRE(k >= 1, VS_ERR_INVALID, "k must be >= 1");
    VS_REQUIRE((int64_t)k <= idx->n_rows, VS_ERR_INVALID, "selected index k out of range (k=%d > N=%lld)", k,
               (long long)idx->n_rows);
    VS_REQUIRE(k <= VS_MAX_K, VS_ERR_UNSUPPORTED, "k=%d > VS_MAX_K=%d", k, VS_MAX_K);
    VS_REQUIRE(mode == VS_MODE_AUTO || mode == VS_MODE_SCAN || mode == VS_MODE_INVERTED, VS_ERR_INVALID, "bad mode");
    VS_REQUIRE(!(d_scores_full && mode == VS_MODE_INVERTED), VS_ERR_INVALID, "vs_scores uses the scan kernels");
    VS_REQUIRE(score_round == VS_F32 || score_round == VS_F16 || score_round == VS_BF16, VS_ERR_INVALID, "bad score_round");
    VS_REQUIRE(workspace_bytes >= vs_search_workspace_bytes(idx, B, k), VS_ERR_INVALID, "workspace too small");
    VS_REQUIRE(idx->n_rows + id_offset < 0xffffffffll && id_offset >= 0, VS_ERR_UNSUPPORTED, "global ids must fit 32 bits");
    if (B == 0) return VS_OK;
    VS_CUDA(cudaSetDevice(idx->device));
    cudaStream_t st = (cudaStream_t)stream;
    const int vpad = vpad_for(idx->n_cols);
    void *ws_base = (void *)(((uintptr_t)d_workspace + 255) / 256 * 256);

    const bool q_on_device = is_device_ptr(hd_q);
    if (idx->kind == 0) {  // dense index: K4 (tcgen05 GEMM + fused top-k)
        VS_REQUIRE(d_scores_full == nullptr, VS_ERR_UNSUPPORTED, "vs_scores is a sparse-path diagnostic");
        Staged sq;
        int rc = sq.init(hd_q, (size_t)B * ldq * dtype_size(q_dtype), st);
        if (rc) return rc;
        rc = search_dense(idx, sq.ptr, q_dtype, B, ldq, k, score_round, id_offset, d_ids, d_scores, d_keys, ws_base, st);
        if (rc == VS_OK && sq.owned) VS_CUDA(cudaStreamSynchronize(st));
        return rc;
    }
    const int64_t chunk = query_chunk(idx, B, k);
    for (int64_t b0 = 0; b0 < B; b0 += chunk) {
        const int64_t Bc = (B - b0) < chunk ? (B - b0) : chunk;
        Workspace w = carve(idx, ws_base, Bc, k);
        const uint8_t *qsrc = (const uint8_t *)hd_q + (size_t)b0 * ldq * dtype_size(q_dtype);
        Staged sq;
        if (!q_on_device) { int rc = sq.init(qsrc, (size_t)Bc * ldq * dtype_size(q_dtype), st); if (rc) return rc; }
        else sq.ptr = qsrc;
        int rc = launch_prep_query(sq.ptr, q_dtype, Bc, ldq, idx->n_cols, vpad, score_round, w.qprep, st);
        if (rc) return rc;
        // ---- scan (K1/K2) or inverted lists (K3)?
        bool use_inv = false;
        uint32_t max_nnz = 0;
        if (mode != VS_MODE_SCAN && !d_scores_full) {
            double mean_post = 0;
            uint64_t max_post = 0;
            rc = inverted_extract(idx, w.qprep, vpad, Bc, w.inv, &max_nnz, &mean_post, &max_post, st);  // SYNC
            if (rc) return rc;
            const bool usable = inverted_usable(max_nnz, max_post);
            if (mode == VS_MODE_INVERTED) {
                VS_REQUIRE(usable, VS_ERR_UNSUPPORTED,
                           "inverted mode needs <= 4096 non-zeros and < 2^32 postings per query (got %u / %llu)", max_nnz,
                           (unsigned long long)max_post);
                use_inv = true;
            } else {
                // cost model (seconds per query), constants measured on B200
                const bool f32 = idx->kind == 1 && idx->store_dtype == VS_F32;
                const double t_scan = (double)idx->stream_bytes / (f32 ? kScanBytesPerSecF32 : kScanBytesPerSecPair);
                const double t_inv = mean_post / kInvPostingsPerSec + (double)idx->n_rows * kInvSecPerRow + kInvFixedSec;
                use_inv = usable && t_inv < t_scan;
            }
        }
        idx->last_mode = use_inv ? VS_MODE_INVERTED : VS_MODE_SCAN;
        const int slot = idx->timer_n < VS_TIMER_SLOTS ? idx->timer_n : -1;
        if (use_inv) {
            rc = launch_inverted(idx, Bc, k, score_round, kInvGroup, max_nnz, w.inv, w.cand,
                                 slot >= 0 ? idx->ev0[slot] : nullptr, slot >= 0 ? idx->ev1[slot] : nullptr, st);
            if (rc) return rc;
        } else {
            if (slot >= 0) VS_CUDA(cudaEventRecord(idx->ev0[slot], st));
            rc = launch_scan(idx, w.qprep, vpad, Bc, k, score_round, w.cand,
                             d_scores_full ? d_scores_full + (size_t)b0 * idx->n_rows : nullptr, st);
            if (rc) return rc;
            if (slot >= 0) VS_CUDA(cudaEventRecord(idx->ev1[slot], st));
        }
        idx->timer_n += 1;
        const int k_in = use_inv ? k : scan_kout(idx, k);   // length of the per-CTA lists
        rc = launch_merge(w.cand, idx->n_ctas, k_in, (int64_t)idx->n_ctas * k_in, Bc, k_in, k, id_offset,
                          d_ids ? d_ids + b0 * k : nullptr, d_scores ? d_scores + b0 * k : nullptr,
                          d_keys ? d_keys + b0 * k : nullptr, st);
        if (rc) return rc;
        if (!q_on_device) VS_CUDA(cudaStreamSynchronize(st));  // staging buffer is freed at scope exit
    }
    return VS_OK;
}

int vs_search(const vs_index *idx, const void *hd_q, int q_dtype, int64_t B, int64_t ldq, int k, int mode,
              int score_round, int64_t id_offset, int64_t *d_ids, float *d_scores, void *d_workspace,
              size_t workspace_bytes, void *stream) {
    VS_REQUIRE(d_ids != nullptr && d_scores != nullptr, VS_ERR_INVALID, "output pointers are NULL");
    return search_impl(idx, hd_q, q_dtype, B, ldq, k, mode, score_round, id_offset, d_ids, d_scores, nullptr, nullptr,
                       d_workspace, workspace_bytes, stream);
}

int vs_search_keys(const vs_index *idx, const void *hd_q, int q_dtype, int64_t B, int64_t ldq, int k, int mode,
                   int score_round, int64_t id_offset, uint64_t *d_keys, void *d_workspace, size_t workspace_bytes,
                   void *stream) {
    VS_REQUIRE(d_keys != nullptr, VS_ERR_INVALID, "output pointer is NULL");
    return search_impl(idx, hd_q, q_dtype, B, ldq, k, mode, score_round, id_offset, nullptr, nullptr, d_keys, nullptr,
                       d_workspace, workspace_bytes, stream);
}

int vs_scores(const vs_index *idx, const void *hd_q, int q_dtype, int64_t B, int64_t ldq, int score_round,
              float *d_scores_full, void *d_workspace, size_t workspace_bytes, void *stream) {
    VS_REQUIRE(d_scores_full != nullptr, VS_ERR_INVALID, "output pointer is NULL");
    return search_impl(idx, hd_q, q_dtype, B, ldq, 1, VS_MODE_SCAN, score_round, 0, nullptr, nullptr, nullptr,
                       d_scores_full, d_workspace, workspace_bytes, stream);
}

int vs_score_rows(const vs_index *idx, const void *hd_q, int q_dtype, int64_t B, int64_t ldq, const int64_t *d_ids, int k,
                  int score_round, float *d_scores, void *d_workspace, size_t workspace_bytes, void *stream) {
    VS_REQUIRE(idx != nullptr && hd_q != nullptr && d_ids != nullptr && d_scores != nullptr, VS_ERR_INVALID, "NULL pointer");
    VS_REQUIRE(idx->kind != 0, VS_ERR_UNSUPPORTED, "vs_score_rows serves sparse and bag-of-token indices");
    VS_REQUIRE(B >= 0 && k >= 1 && ldq >= idx->n_cols, VS_ERR_INVALID, "bad query / candidate shape");
    VS_REQUIRE(q_dtype == VS_F32 || q_dtype == VS_F16 || q_dtype == VS_BF16, VS_ERR_INVALID, "queries must be f32 / f16 / bf16");
    if (B == 0) return VS_OK;
    VS_CUDA(cudaSetDevice(idx->device));
    cudaStream_t st = (cudaStream_t)stream;
    const int vpad = vpad_for(idx->n_cols);
    void *ws = (void *)(((uintptr_t)d_workspace + 255) / 256 * 256);
    VS_REQUIRE(d_workspace != nullptr && workspace_bytes >= (size_t)B * vpad * 4 + 256, VS_ERR_INVALID,
               "workspace too small: vs_score_rows needs B * %d * 4 + 256 bytes", vpad);
    Staged sq;
    int rc = sq.init(hd_q, (size_t)B * ldq * dtype_size(q_dtype), st);
    if (rc) return rc;
    rc = launch_prep_query(sq.ptr, q_dtype, B, ldq, idx->n_cols, vpad, score_round, (float *)ws, st);
    if (rc) return rc;
    const int64_t n_pairs = B * (int64_t)k;
    score_rows_kernel<<<(unsigned)((n_pairs + 7) / 8), 256, 0, st>>>(ws_view(idx), (const float *)ws, vpad, d_ids, n_pairs, k,
                                                                    score_round, d_scores);
    VS_CUDA(cudaGetLastError());
    if (sq.owned) VS_CUDA(cudaStreamSynchronize(st));   // staging buffer is freed at scope exit
    return VS_OK;
}

int vs_merge_keys(int device, const uint64_t *d_keys_in, int64_t P, int64_t stride_p, int64_t stride_b, int64_t B,
                  int k_in, int k_out, int64_t *d_ids, float *d_scores, void *stream) {
    VS_REQUIRE(d_keys_in != nullptr && d_ids != nullptr && d_scores != nullptr, VS_ERR_INVALID, "NULL pointer");
    VS_REQUIRE(P >= 1 && k_in >= 1 && k_out >= 1 && B >= 0, VS_ERR_INVALID, "bad merge shape");
    VS_CUDA(cudaSetDevice(device));
    return launch_merge(d_keys_in, P, stride_p, stride_b, B, k_in, k_out, 0, d_ids, d_scores, nullptr,
                        (cudaStream_t)stream);
}

int vs_index_last_mode(const vs_index *idx, int *mode) {
    VS_REQUIRE(idx != nullptr && mode != nullptr, VS_ERR_INVALID, "NULL pointer");
    *mode = idx->last_mode;
    return VS_OK;
}

int vs_debug_gather_wavefronts(const vs_index *idx, unsigned long long *d_out2, void *stream) {
    VS_REQUIRE(idx != nullptr && idx->kind != 0 && d_out2 != nullptr, VS_ERR_INVALID, "needs a sparse / binary index");
    VS_CUDA(cudaSetDevice(idx->device));
    return debug_gather_wavefronts(idx, d_out2, (cudaStream_t)stream);
}

int vs_debug_scan_profile(vs_index *idx, unsigned long long *d_buf) {
    VS_REQUIRE(idx != nullptr, VS_ERR_INVALID, "NULL pointer");
    idx->scan_prof = d_buf;
    return VS_OK;
}

int vs_kernel_timer(vs_index *idx, int reset, float *total_ms, int *launches) {
    VS_REQUIRE(idx != nullptr, VS_ERR_INVALID, "NULL pointer");
    VS_CUDA(cudaSetDevice(idx->device));
    const int n = idx->timer_n < VS_TIMER_SLOTS ? idx->timer_n : VS_TIMER_SLOTS;
    float total = 0.f;
    for (int i = 0; i < n; ++i) {
        float ms = 0.f;
        VS_CUDA(cudaEventSynchronize(idx->ev1[i]));
        VS_CUDA(cudaEventElapsedTime(&ms, idx->ev0[i], idx->ev1[i]));
        total += ms;
    }
    if (total_ms) *total_ms = total;
    if (launches) *launches = n;
    if (reset) idx->timer_n = 0;
    return VS_OK;
}

}  // extern "C"
