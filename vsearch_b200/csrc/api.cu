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

// scan.cu / merge.cu / inverted.cu / dense.cu
size_t scan_smem_bytes(int vpad, int cap);
int scan_cap_for_k(int k);
int scan_kout(const vs_index *idx, int k);
int launch_prep_query(const void *d_q, int q_dtype, int64_t B, int64_t ldq, int64_t n_cols, int vpad, int round_mode,
                      float *d_out, cudaStream_t st);
int launch_scan(const vs_index *idx, const float *d_qprep, const SparseQueries *sq, int vpad, int64_t B, int k,
                int score_round, uint64_t *d_cand, float *d_scores_out, const int *d_use_inv, cudaStream_t st);
size_t inverted_workspace_bytes(const vs_index *idx, int64_t Bc, int group);
int inverted_prepare(vs_index *idx, const float *d_qprep, int vpad, const void *d_qptr, int ptr_dtype, const int32_t *d_qtok,
                     const float *d_qw, int64_t b0, int64_t Bc, int mode, int round_mode, double t_scan, double postings_per_sec,
                     double t_rows_fixed, void *d_ws, int *d_flag, cudaStream_t st);
int launch_inverted(vs_index *idx, int64_t Bc, int k, int cand_stride, int score_round, void *d_ws, uint64_t *d_cand,
                    const int *d_flag, cudaStream_t st);
size_t dense_workspace_bytes(const vs_index *idx, int64_t B, int k);
int search_dense_step(vs_index *idx, int step, const void *d_q, int q_dtype, int64_t B, int64_t ldq, int k, int score_round,
                      int64_t id_offset, int n_ranks, const uint64_t *d_gathered, uint64_t *d_keys_out, uint32_t *d_status,
                      void *d_ws, cudaStream_t st);
int search_dense(vs_index *idx, const void *d_q, int q_dtype, int64_t B, int64_t ldq, int k, int score_round,
                 int64_t id_offset, int64_t *d_ids, float *d_scores, uint64_t *d_keys, void *d_ws, cudaStream_t st);
int launch_merge(const uint64_t *d_in, int64_t P, int64_t stride_p, int64_t stride_b, int64_t B, int k_in, int k_out,
                 int64_t id_offset, int64_t *d_ids, float *d_scores, uint64_t *d_keys, cudaStream_t st);
size_t merge_scratch_bytes(int64_t P, int k_in, int k_out, int64_t B);
int launch_merge_staged(const uint64_t *d_in, int64_t P, int64_t stride_p, int64_t stride_b, int64_t B, int k_in, int k_out,
                        int64_t id_offset, int64_t *d_ids, float *d_scores, uint64_t *d_keys, const int *d_alt_flag, int k_in_alt,
                        void *d_scratch, cudaStream_t st);

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
    int *flag;         // device decision of the auto mode (1 = inverted lists)
    float *qprep;      // [Bc, vpad] prepared fp32 rows
    uint64_t *cand;    // [Bc, n_ctas, scan_kout] per-CTA candidate lists
    void *inv;         // K3: extracted (token, weight) lists
    uint8_t *stage;    // host inputs land here first: dense rows [Bc, n_cols] in the caller's dtype, or sparse lists
    size_t stage_bytes;
    void *merge;       // group results of the multi-stage merge (merge.cu)
    size_t bytes;
};
constexpr int kInvGroup = 1;   // queries scored concurrently by the inverted-list path (accumulator rows)
static size_t align256(size_t x) { return (x + 255) / 256 * 256; }
static Workspace carve(const vs_index *idx, void *base, int64_t Bc, int k) {
    Workspace w;
    const size_t q_bytes = align256((size_t)Bc * vpad_for(idx->n_cols) * 4);
    const size_t c_bytes = align256((size_t)Bc * idx->n_ctas * (size_t)scan_kout(idx, k) * 8);   // scan lists are the longer ones
    const size_t i_bytes = align256(inverted_workspace_bytes(idx, Bc, kInvGroup));
    // staging: a dense fp32 chunk, or a sparse chunk as dense as the rows themselves (8 B per entry + offsets)
    w.stage_bytes = align256((size_t)Bc * (size_t)idx->n_cols * 8 + (size_t)(Bc + 1) * 8);
    uint8_t *p = (uint8_t *)base;
    w.flag = (int *)p; p += 256;
    w.qprep = (float *)p; p += q_bytes;
    w.cand = (uint64_t *)p; p += c_bytes;
    w.inv = p; p += i_bytes;
    w.stage = p; p += w.stage_bytes;
    w.merge = p; p += align256(merge_scratch_bytes(idx->n_ctas, scan_kout(idx, k), k, Bc));
    w.bytes = (size_t)(p - (uint8_t *)base);
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
// auto-mode cost model, fitted to profiles/r2n_sweep_*.jsonl (21M rows, 32..768 tokens per query, 1 GPU): a query through
// the inverted lists costs postings / rate + rows * kInvSecPerRow + kInvFixedSec (36 us at 64 tokens on the binary
// config-2 index, 69 us on the fp32 config-5 index), through the scan one pass over the stream
constexpr double kScanBytesPerSecPair = 5.6e12;   // binary / 16-bit values
constexpr double kScanBytesPerSecF32 = 6.2e12;    // fp32 values (HBM bound)
constexpr double kInvPostingsPerSecBinary = 3.7e11;   // fixed-point integer adds
constexpr double kInvPostingsPerSecValued = 2.5e11;   // fp32 CAS loops + the value loads
constexpr double kInvSecPerRow = 0.8e-12;         // zero + select of the block accumulators
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

int create_csr_from_device(int device, int64_t n_rows, int64_t n_cols, int64_t nnz, const void *d_crow, int crow_dtype,
                           const void *d_col, int col_dtype, const void *d_val, int val_dtype, int store_dtype,
                           int64_t col_shift, cudaStream_t st, vs_index **out) {
    *out = nullptr;
    VS_REQUIRE(n_rows >= 0 && n_cols >= 1 && nnz >= 0 && col_shift >= 0, VS_ERR_INVALID, "bad shape");
    VS_REQUIRE(n_cols <= 32767, VS_ERR_UNSUPPORTED,
               "n_cols=%lld does not fit the uint16 column format (15 bits + the row-end flag)", (long long)n_cols);
    VS_REQUIRE(n_rows < 0xffffffffll, VS_ERR_UNSUPPORTED, "n_rows must be < 2^32 - 1 per shard");
    const bool binary = (d_val == nullptr || val_dtype == VS_NONE);
    if (!binary) {
        VS_REQUIRE(val_dtype == VS_F32 || val_dtype == VS_F16 || val_dtype == VS_BF16, VS_ERR_INVALID, "bad value dtype");
        VS_REQUIRE(store_dtype == VS_F32 || store_dtype == VS_F16 || store_dtype == VS_BF16, VS_ERR_INVALID, "bad store dtype");
    }
    VS_CUDA(cudaSetDevice(device));
    vs_index *idx = new (std::nothrow) vs_index();
    VS_REQUIRE(idx != nullptr, VS_ERR_NOMEM, "out of host memory");
    idx->device = device;
    idx->kind = binary ? 2 : 1;
    idx->store_dtype = binary ? VS_NONE : store_dtype;
    idx->n_rows = n_rows; idx->n_cols = n_cols; idx->nnz = nnz;
    if (const char *e = getenv("VSEARCH_B200_BANK_AWARE")) idx->bank_aware = (e[0] != '0');
    if (cudaMalloc(&idx->d_last_mode, 256) != cudaSuccess) { delete idx; VS_REQUIRE(false, VS_ERR_NOMEM, "out of device memory"); }
    cudaMemsetAsync(idx->d_last_mode, 0, 256, st);
    int rc = build_ws_index(idx, d_crow, crow_dtype, d_col, col_dtype, binary ? nullptr : d_val, binary ? VS_NONE : val_dtype, st,
                            col_shift);
    cudaStreamSynchronize(st);
    for (int i = 0; i < VS_TIMER_SLOTS && rc == VS_OK; ++i)
        if (cudaEventCreate(&idx->ev0[i]) != cudaSuccess || cudaEventCreate(&idx->ev1[i]) != cudaSuccess) {
            set_error("cudaEventCreate failed");
            rc = VS_ERR_CUDA;
        }
    if (rc != VS_OK) { vs_index_destroy(idx); return rc; }
    *out = idx;
    return VS_OK;
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
    VS_REQUIRE(crow_dtype == VS_I32 || crow_dtype == VS_I64, VS_ERR_INVALID, "crow dtype must be int32/int64");
    VS_REQUIRE(col_dtype == VS_I32 || col_dtype == VS_I64, VS_ERR_INVALID, "col dtype must be int32/int64");
    const bool binary = (hd_val == nullptr || val_dtype == VS_NONE);
    VS_REQUIRE(hd_crow != nullptr && (nnz == 0 || hd_col != nullptr), VS_ERR_INVALID, "NULL CSR arrays");
    VS_CUDA(cudaSetDevice(device));
    cudaStream_t st = (cudaStream_t)stream;
    Staged s_crow, s_col, s_val;
    int rc = s_crow.init(hd_crow, (size_t)(n_rows + 1) * dtype_size(crow_dtype), st);
    if (rc == VS_OK) rc = s_col.init(hd_col, (size_t)nnz * dtype_size(col_dtype), st);
    if (rc == VS_OK && !binary) rc = s_val.init(hd_val, (size_t)nnz * dtype_size(val_dtype), st);
    if (rc == VS_OK)
        rc = create_csr_from_device(device, n_rows, n_cols, nnz, s_crow.ptr, crow_dtype, s_col.ptr, col_dtype, binary ? nullptr : s_val.ptr,
                                    binary ? VS_NONE : val_dtype, store_dtype, 0, st, out);
    cudaStreamSynchronize(st);   // the staging copies are freed at scope exit
    return rc;
}

int vs_index_create_dense(int device, int64_t n_rows, int64_t dim, const void *hd_x, int x_dtype, int64_t ld,
                          int store_dtype, void *stream, vs_index **out) {
    VS_REQUIRE(out != nullptr, VS_ERR_INVALID, "out is NULL");
    *out = nullptr;
    VS_REQUIRE(n_rows >= 0 && dim >= 1 && ld >= dim && hd_x != nullptr, VS_ERR_INVALID, "bad dense shape");
    VS_REQUIRE(x_dtype == VS_F32 || x_dtype == VS_F16 || x_dtype == VS_BF16, VS_ERR_INVALID, "bad dense dtype");
    VS_REQUIRE(store_dtype == VS_F16 || store_dtype == VS_BF16 || store_dtype == VS_F32, VS_ERR_INVALID, "bad dense store dtype");
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
    int prev = 0;
    cudaGetDevice(&prev);   // the caller's current device is left as it was (the handle may live on another GPU)
    cudaSetDevice(idx->device);
    cudaFree(idx->d_last_mode);
    cudaFree(idx->cols); cudaFree(idx->vals); cudaFree(idx->tails);
    cudaFree(idx->part_win_begin); cudaFree(idx->part_row_begin); cudaFree(idx->row_chunk);
    cudaFree(idx->dense); cudaFree(idx->dense32);
    cudaFree(idx->post_ptr); cudaFree(idx->blk_ptr); cudaFree(idx->blk_base); cudaFree(idx->post_row); cudaFree(idx->post_val);
    for (int i = 0; i < VS_TIMER_SLOTS; ++i) {
        if (idx->ev0[i]) cudaEventDestroy(idx->ev0[i]);
        if (idx->ev1[i]) cudaEventDestroy(idx->ev1[i]);
    }
    delete idx;
    cudaSetDevice(prev);
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
    if (idx->kind == 0) return dense_workspace_bytes(idx, B, k) + align256((size_t)B * (size_t)idx->dim * 4) + 1024 + 256;
    return carve(idx, nullptr, query_chunk(idx, B, k), k).bytes + 256;
}

// Queries of one call: dense rows (q != nullptr) or CSR-style (token, weight) lists (ptr != nullptr); host or device.
struct QueryInput {
    const void *q = nullptr; int q_dtype = VS_F32; int64_t ldq = 0;
    const void *ptr = nullptr; int ptr_dtype = VS_I64; const int32_t *tok = nullptr; const float *w = nullptr;
};

static int search_impl(const vs_index *cidx, const QueryInput &in, int64_t B, int k, int mode, int score_round,
                       int64_t id_offset, int64_t *d_ids, float *d_scores, uint64_t *d_keys, float *d_scores_full,
                       void *d_workspace, size_t workspace_bytes, void *stream) {
    vs_index *idx = const_cast<vs_index *>(cidx);
    VS_REQUIRE(idx != nullptr, VS_ERR_INVALID, "index is NULL");
    const bool sparse_q = in.ptr != nullptr;
    if (!sparse_q) {
        VS_REQUIRE(B >= 0 && in.ldq >= idx->n_cols, VS_ERR_INVALID, "query leading dimension %lld < n_cols %lld", (long long)in.ldq,
                   (long long)idx->n_cols);
        VS_REQUIRE(in.q_dtype == VS_F32 || in.q_dtype == VS_F16 || in.q_dtype == VS_BF16, VS_ERR_INVALID, "bad query dtype");
        VS_REQUIRE(in.q != nullptr || B == 0, VS_ERR_INVALID, "queries are NULL");
    } else {
        VS_REQUIRE(idx->kind != 0, VS_ERR_UNSUPPORTED, "sparse queries need a sparse / bag-of-token index");
        VS_REQUIRE(in.ptr_dtype == VS_I32 || in.ptr_dtype == VS_I64, VS_ERR_INVALID, "query offsets must be int32 / int64");
        VS_REQUIRE(B >= 0, VS_ERR_INVALID, "bad batch size");
    }
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
    uint8_t *ws_base = (uint8_t *)(((uintptr_t)d_workspace + 255) / 256 * 256);

    if (idx->kind == 0) {  // dense index: K4 (tcgen05 GEMM + fused top-k).  SYNC once, at the end (overflow status word)
        VS_REQUIRE(d_scores_full == nullptr, VS_ERR_UNSUPPORTED, "vs_scores is a sparse-path diagnostic");
        const void *dq = in.q;
        size_t ws_off = 0;
        int64_t ld = in.ldq;
        if (!is_device_ptr(in.q)) {   // stage the host batch (compact [B, dim]) at the head of the workspace
            const size_t esz = dtype_size(in.q_dtype);
            VS_CUDA(cudaMemcpy2DAsync(ws_base, (size_t)idx->dim * esz, in.q, (size_t)in.ldq * esz, (size_t)idx->dim * esz, (size_t)B,
                                      cudaMemcpyHostToDevice, st));
            dq = ws_base;
            ld = idx->dim;
            ws_off = align256((size_t)B * idx->dim * esz);
        }
        idx->last_mode = VS_MODE_SCAN; idx->last_mode_on_device = false;
        return search_dense(idx, dq, in.q_dtype, B, ld, k, score_round, id_offset, d_ids, d_scores, d_keys, ws_base + ws_off, st);
    }

    const bool q_on_device = sparse_q ? is_device_ptr(in.ptr) : is_device_ptr(in.q);
    const bool f32 = idx->kind == 1 && idx->store_dtype == VS_F32;
    const double t_scan = (double)idx->stream_bytes / (f32 ? kScanBytesPerSecF32 : kScanBytesPerSecPair);
    const double t_rows_fixed = (double)idx->n_rows * kInvSecPerRow + kInvFixedSec;
    const int kout = scan_kout(idx, k);
    const bool try_inv = mode != VS_MODE_SCAN && !d_scores_full;
    const int64_t chunk = query_chunk(idx, B, k);
    for (int64_t b0 = 0; b0 < B; b0 += chunk) {
        const int64_t Bc = (B - b0) < chunk ? (B - b0) : chunk;
        Workspace w = carve(idx, ws_base, Bc, k);
        SparseQueries sq;
        int rc;
        if (!sparse_q) {
            // ---- dense rows -> prepared fp32 [Bc, vpad]
            const size_t esz = dtype_size(in.q_dtype);
            const uint8_t *qsrc = (const uint8_t *)in.q + (size_t)b0 * in.ldq * esz;
            int64_t ld = in.ldq;
            if (!q_on_device) {   // compact [Bc, n_cols] copy into the staging area (no allocation, no host wait)
                VS_CUDA(cudaMemcpy2DAsync(w.stage, (size_t)idx->n_cols * esz, qsrc, (size_t)in.ldq * esz, (size_t)idx->n_cols * esz,
                                          (size_t)Bc, cudaMemcpyHostToDevice, st));
                qsrc = w.stage;
                ld = idx->n_cols;
            }
            rc = launch_prep_query(qsrc, in.q_dtype, Bc, ld, idx->n_cols, vpad, score_round, w.qprep, st);
            if (rc) return rc;
        } else {
            // ---- (token, weight) lists: used as they are (device) or staged (host: offsets first, then the entries)
            sq.ptr = in.ptr; sq.ptr_dtype = in.ptr_dtype; sq.tok = in.tok; sq.w = in.w; sq.b0 = b0;
            if (!q_on_device) {
                const size_t psz = dtype_size(in.ptr_dtype);
                const uint8_t *hp = (const uint8_t *)in.ptr + (size_t)b0 * psz;
                const int64_t lo = in.ptr_dtype == VS_I32 ? (int64_t)((const int32_t *)in.ptr)[b0] : ((const int64_t *)in.ptr)[b0];
                const int64_t hi = in.ptr_dtype == VS_I32 ? (int64_t)((const int32_t *)in.ptr)[b0 + Bc] : ((const int64_t *)in.ptr)[b0 + Bc];
                VS_REQUIRE(lo >= 0 && hi >= lo, VS_ERR_INVALID, "query offsets must be non-decreasing");
                const size_t n_ent = (size_t)(hi - lo);
                const size_t p_bytes = align256((size_t)(Bc + 1) * psz);
                VS_REQUIRE(p_bytes + 2 * align256(n_ent * 4) <= w.stage_bytes, VS_ERR_UNSUPPORTED,
                           "sparse query chunk holds more entries than dense rows would");
                uint8_t *d_ptr = w.stage, *d_tok = w.stage + p_bytes, *d_w = d_tok + align256(n_ent * 4);
                VS_CUDA(cudaMemcpyAsync(d_ptr, hp, (size_t)(Bc + 1) * psz, cudaMemcpyHostToDevice, st));
                if (n_ent) {
                    VS_CUDA(cudaMemcpyAsync(d_tok, in.tok + lo, n_ent * 4, cudaMemcpyHostToDevice, st));
                    VS_CUDA(cudaMemcpyAsync(d_w, in.w + lo, n_ent * 4, cudaMemcpyHostToDevice, st));
                }
                // the staged offsets still count from the caller's array start: rebase the entry pointers instead
                sq.ptr = d_ptr; sq.tok = (const int32_t *)d_tok - lo; sq.w = (const float *)d_w - lo; sq.b0 = 0;
            }
        }
        // ---- scan (K1/K2) or inverted lists (K3)?  Decided on the device (inv_decide_kernel): both kernels are
        // enqueued, the one that lost exits at once, the merge reads the same flag.  No readback, no host wait.
        if (try_inv) {
            rc = inverted_prepare(idx, w.qprep, vpad, sparse_q ? sq.ptr : nullptr, sq.ptr_dtype, sparse_q ? sq.tok : nullptr,
                                  sparse_q ? sq.w : nullptr, sparse_q ? sq.b0 : 0, Bc, mode, score_round, t_scan,
                                  idx->kind == 1 ? kInvPostingsPerSecValued : kInvPostingsPerSecBinary, t_rows_fixed, w.inv, w.flag, st);
            if (rc) return rc;
            idx->last_mode_on_device = true;
        } else {
            idx->last_mode = VS_MODE_SCAN; idx->last_mode_on_device = false;
        }
        const int slot = idx->timer_n < VS_TIMER_SLOTS ? idx->timer_n : -1;
        if (slot >= 0) VS_CUDA(cudaEventRecord(idx->ev0[slot], st));
        if (try_inv) {
            rc = launch_inverted(idx, Bc, k, kout, score_round, w.inv, w.cand, w.flag, st);
            if (rc) return rc;
        }
        rc = launch_scan(idx, w.qprep, sparse_q ? &sq : nullptr, vpad, Bc, k, score_round, w.cand,
                         d_scores_full ? d_scores_full + (size_t)b0 * idx->n_rows : nullptr, try_inv ? w.flag : nullptr, st);
        if (rc) return rc;
        if (slot >= 0) VS_CUDA(cudaEventRecord(idx->ev1[slot], st));
        idx->timer_n += 1;
        // per-CTA lists: up to scan_kout keys each (zero padded), whichever kernel family wrote them
        rc = launch_merge_staged(w.cand, idx->n_ctas, kout, (int64_t)idx->n_ctas * kout, Bc, kout, k, id_offset,
                                 d_ids ? d_ids + b0 * k : nullptr, d_scores ? d_scores + b0 * k : nullptr,
                                 d_keys ? d_keys + b0 * k : nullptr, nullptr, 0, w.merge, st);
        if (rc) return rc;
    }
    return VS_OK;
}

int vs_search(const vs_index *idx, const void *hd_q, int q_dtype, int64_t B, int64_t ldq, int k, int mode,
              int score_round, int64_t id_offset, int64_t *d_ids, float *d_scores, void *d_workspace,
              size_t workspace_bytes, void *stream) {
    VS_REQUIRE(d_ids != nullptr && d_scores != nullptr, VS_ERR_INVALID, "output pointers are NULL");
    QueryInput in;
    in.q = hd_q; in.q_dtype = q_dtype; in.ldq = ldq;
    return search_impl(idx, in, B, k, mode, score_round, id_offset, d_ids, d_scores, nullptr, nullptr, d_workspace,
                       workspace_bytes, stream);
}

int vs_search_keys(const vs_index *idx, const void *hd_q, int q_dtype, int64_t B, int64_t ldq, int k, int mode,
                   int score_round, int64_t id_offset, uint64_t *d_keys, void *d_workspace, size_t workspace_bytes,
                   void *stream) {
    VS_REQUIRE(d_keys != nullptr, VS_ERR_INVALID, "output pointer is NULL");
    QueryInput in;
    in.q = hd_q; in.q_dtype = q_dtype; in.ldq = ldq;
    return search_impl(idx, in, B, k, mode, score_round, id_offset, nullptr, nullptr, d_keys, nullptr, d_workspace,
                       workspace_bytes, stream);
}

int vs_search_dense_step(const vs_index *cidx, int step, const void *d_q, int q_dtype, int64_t B, int64_t ldq, int k,
                         int score_round, int64_t id_offset, int n_ranks, const uint64_t *d_gathered_keys, uint64_t *d_keys,
                         uint32_t *d_status, void *d_workspace, size_t workspace_bytes, void *stream) {
    vs_index *idx = const_cast<vs_index *>(cidx);
    VS_REQUIRE(idx != nullptr, VS_ERR_INVALID, "index is NULL");
    VS_REQUIRE(idx->kind == 0, VS_ERR_UNSUPPORTED, "vs_search_dense_step needs a dense index");
    VS_REQUIRE(d_q != nullptr && d_keys != nullptr, VS_ERR_INVALID, "queries / output keys are NULL");
    VS_REQUIRE(q_dtype == VS_F32 || q_dtype == VS_F16 || q_dtype == VS_BF16, VS_ERR_INVALID, "bad query dtype");
    VS_REQUIRE(ldq >= idx->n_cols, VS_ERR_INVALID, "query leading dimension %lld < dim %lld", (long long)ldq, (long long)idx->n_cols);
    VS_REQUIRE(k >= 1 && (int64_t)k <= idx->n_rows && k <= VS_MAX_K, VS_ERR_INVALID, "k=%d out of range (rows %lld, VS_MAX_K %d)", k,
               (long long)idx->n_rows, VS_MAX_K);
    VS_REQUIRE(score_round == VS_F32 || score_round == VS_F16 || score_round == VS_BF16, VS_ERR_INVALID, "bad score_round");
    VS_REQUIRE(B >= 1 && workspace_bytes >= vs_search_workspace_bytes(idx, B, k), VS_ERR_INVALID, "bad batch / workspace too small");
    VS_REQUIRE(is_device_ptr(d_q), VS_ERR_INVALID, "vs_search_dense_step takes device queries");
    VS_CUDA(cudaSetDevice(idx->device));
    uint8_t *ws_base = (uint8_t *)(((uintptr_t)d_workspace + 255) / 256 * 256);
    idx->last_mode = VS_MODE_SCAN; idx->last_mode_on_device = false;
    return search_dense_step(idx, step, d_q, q_dtype, B, ldq, k, score_round, id_offset, n_ranks, d_gathered_keys, d_keys, d_status,
                             ws_base, (cudaStream_t)stream);
}

int vs_search_sparse(const vs_index *idx, const void *hd_qptr, int ptr_dtype, const int32_t *hd_qtok, const float *hd_qw,
                     int64_t B, int k, int mode, int score_round, int64_t id_offset, int64_t *d_ids, float *d_scores,
                     uint64_t *d_keys, void *d_workspace, size_t workspace_bytes, void *stream) {
    VS_REQUIRE((d_ids != nullptr && d_scores != nullptr) || d_keys != nullptr, VS_ERR_INVALID, "output pointers are NULL");
    VS_REQUIRE(hd_qptr != nullptr, VS_ERR_INVALID, "query offsets are NULL");
    QueryInput in;
    in.ptr = hd_qptr; in.ptr_dtype = ptr_dtype; in.tok = hd_qtok; in.w = hd_qw;
    return search_impl(idx, in, B, k, mode, score_round, id_offset, d_ids, d_scores, d_keys, nullptr, d_workspace,
                       workspace_bytes, stream);
}

int vs_scores(const vs_index *idx, const void *hd_q, int q_dtype, int64_t B, int64_t ldq, int score_round,
              float *d_scores_full, void *d_workspace, size_t workspace_bytes, void *stream) {
    VS_REQUIRE(d_scores_full != nullptr, VS_ERR_INVALID, "output pointer is NULL");
    QueryInput in;
    in.q = hd_q; in.q_dtype = q_dtype; in.ldq = ldq;
    return search_impl(idx, in, B, 1, VS_MODE_SCAN, score_round, 0, nullptr, nullptr, nullptr, d_scores_full, d_workspace,
                       workspace_bytes, stream);
}

size_t vs_score_rows_workspace_bytes(const vs_index *idx, int64_t B) {
    if (!idx || B <= 0) return 256;
    return align256((size_t)B * vpad_for(idx->n_cols) * 4) + align256((size_t)B * (size_t)idx->n_cols * 4) + 512;
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
    VS_REQUIRE(d_workspace != nullptr && workspace_bytes >= vs_score_rows_workspace_bytes(idx, B), VS_ERR_INVALID,
               "workspace too small: vs_score_rows needs vs_score_rows_workspace_bytes(idx, B) bytes");
    const void *dq = hd_q;
    int64_t ld = ldq;
    if (!is_device_ptr(hd_q)) {   // host batch: compact copy behind the prepared rows, no allocation, no host wait
        uint8_t *stage = (uint8_t *)ws + align256((size_t)B * vpad * 4);
        const size_t esz = dtype_size(q_dtype);
        VS_CUDA(cudaMemcpy2DAsync(stage, (size_t)idx->n_cols * esz, hd_q, (size_t)ldq * esz, (size_t)idx->n_cols * esz, (size_t)B,
                                  cudaMemcpyHostToDevice, st));
        dq = stage;
        ld = idx->n_cols;
    }
    int rc = launch_prep_query(dq, q_dtype, B, ld, idx->n_cols, vpad, score_round, (float *)ws, st);
    if (rc) return rc;
    const int64_t n_pairs = B * (int64_t)k;
    score_rows_kernel<<<(unsigned)((n_pairs + 7) / 8), 256, 0, st>>>(ws_view(idx), (const float *)ws, vpad, d_ids, n_pairs, k,
                                                                    score_round, d_scores);
    VS_CUDA(cudaGetLastError());
    return VS_OK;
}

int vs_merge_keys(int device, const uint64_t *d_keys_in, int64_t P, int64_t stride_p, int64_t stride_b, int64_t B,
                  int k_in, int k_out, int64_t *d_ids, float *d_scores, void *stream) {
    VS_REQUIRE(d_keys_in != nullptr && d_ids != nullptr && d_scores != nullptr, VS_ERR_INVALID, "NULL pointer");
    VS_REQUIRE(P >= 1 && k_in >= 1 && k_out >= 1 && B >= 0, VS_ERR_INVALID, "bad merge shape");
    VS_CUDA(cudaSetDevice(device));
    // register-resident kernel when the P lists fit one CTA (8 ranks x k <= 2048 do), else the streaming one
    return launch_merge_staged(d_keys_in, P, stride_p, stride_b, B, k_in, k_out, 0, d_ids, d_scores, nullptr, nullptr, 0, nullptr,
                               (cudaStream_t)stream);
}

int vs_index_last_mode(const vs_index *idx, int *mode) {
    VS_REQUIRE(idx != nullptr && mode != nullptr, VS_ERR_INVALID, "NULL pointer");
    *mode = idx->last_mode;
    if (idx->last_mode_on_device) {   // the device chose (auto / inverted): read its decision back.  SYNC
        int prev = 0;
        cudaGetDevice(&prev);
        VS_CUDA(cudaSetDevice(idx->device));
        cudaError_t e = cudaMemcpy(mode, idx->d_last_mode, sizeof(int), cudaMemcpyDeviceToHost);
        cudaSetDevice(prev);
        VS_CUDA(e);
    }
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
    int prev = 0;
    cudaGetDevice(&prev);
    VS_CUDA(cudaSetDevice(idx->device));
    const int n = idx->timer_n < VS_TIMER_SLOTS ? idx->timer_n : VS_TIMER_SLOTS;
    float total = 0.f;
    cudaError_t err = cudaSuccess;
    for (int i = 0; i < n && err == cudaSuccess; ++i) {
        float ms = 0.f;
        err = cudaEventSynchronize(idx->ev1[i]);
        if (err == cudaSuccess) err = cudaEventElapsedTime(&ms, idx->ev0[i], idx->ev1[i]);
        total += ms;
    }
    cudaSetDevice(prev);   // leave the caller's current device as it was
    VS_CUDA(err);
    if (total_ms) *total_ms = total;
    if (launches) *launches = n;
    if (reset) idx->timer_n = 0;
    return VS_OK;
}

}  // extern "C"
