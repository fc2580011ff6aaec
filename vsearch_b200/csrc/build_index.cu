// build_index.cu -- CSR triple -> warp-stream (WS) device format, and back.
// Replaces SparseIndex._scipy_csr_to_torch_csr + .to(device) (upstream index.py:144-161,179).
// One-off work at index load; not on the search path (CUB's scan is used for the prefix sum).
#include <cub/device/device_scan.cuh>

#include "index.cuh"

namespace vs {

template <typename T>
__device__ __forceinline__ int64_t load_idx(const void *p, int64_t i) { return (int64_t)((const T *)p)[i]; }
__device__ __forceinline__ int64_t load_index(const void *p, int dtype, int64_t i) {
    return dtype == VS_I32 ? load_idx<int32_t>(p, i) : load_idx<int64_t>(p, i);
}
__device__ __forceinline__ float load_value(const void *p, int dtype, int64_t i) {
    if (dtype == VS_F32) return ((const float *)p)[i];
    if (dtype == VS_F16) return __half2float(((const __half *)p)[i]);
    return __bfloat162float(((const __nv_bfloat16 *)p)[i]);
}

// chunks per row (>= 1 so that every row, even an empty one, owns a tail bit)
__global__ void row_chunks_kernel(const void *crow, int crow_dtype, int64_t n_rows, uint64_t *rc, int *err) {
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r > n_rows) return;
    if (r == n_rows) { rc[r] = 0; return; }
    int64_t a = load_index(crow, crow_dtype, r), e = load_index(crow, crow_dtype, r + 1);
    if (e < a) { atomicExch(err, 1); e = a; }
    uint64_t len = (uint64_t)(e - a);
    rc[r] = len == 0 ? 1 : (len + 7) / 8;
}

// part p starts at the first row whose chunk offset is >= p/n_parts of the total
__global__ void part_rows_kernel(const uint64_t *cptr, int64_t n_rows, int n_parts, uint32_t *part_row_begin) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p > n_parts) return;
    if (p == n_parts) { part_row_begin[p] = (uint32_t)n_rows; return; }
    uint64_t total = cptr[n_rows];
    uint64_t target = (total / (uint64_t)n_parts) * (uint64_t)p + (total % (uint64_t)n_parts) * (uint64_t)p / (uint64_t)n_parts;
    int64_t lo = 0, hi = n_rows;  // first r in [0, N] with cptr[r] >= target
    while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        if (cptr[mid] >= target) hi = mid; else lo = mid + 1;
    }
    part_row_begin[p] = (uint32_t)lo;
}

__global__ void part_windows_kernel(const uint64_t *cptr, const uint32_t *part_row_begin, int n_parts,
                                    uint32_t *part_win_begin, uint64_t *n_windows_out) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    uint64_t w = 0;
    for (int p = 0; p < n_parts; ++p) {
        part_win_begin[p] = (uint32_t)w;
        uint64_t chunks = cptr[part_row_begin[p + 1]] - cptr[part_row_begin[p]];
        w += (chunks + 31) / 32;
    }
    part_win_begin[n_parts] = (uint32_t)w;
    *n_windows_out = w;
}

__global__ void fill_u32_kernel(uint32_t *p, uint64_t n, uint32_t v) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) p[i] = v;
}

// one warp per row: scatter the row's entries into its chunks, set the tail bit, record row_chunk
template <typename VT>
__global__ void fill_rows_kernel(const void *crow, int crow_dtype, const void *col, int col_dtype, const void *val,
                                 int val_dtype, int64_t n_rows, int64_t n_cols, const uint64_t *cptr,
                                 const uint32_t *part_row_begin, const uint32_t *part_win_begin, int n_parts,
                                 uint16_t *cols16, VT *vals, uint32_t *tails, uint32_t *row_chunk, int *err) {
    int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (r >= n_rows) return;
    // last part p with part_row_begin[p] <= r  (empty parts repeat the same begin; the last one owns r)
    int lo = 0, hi = n_parts;  // invariant: part_row_begin[lo] <= r
    while (lo < hi) {
        int mid = (lo + hi + 1) >> 1;
        if ((int64_t)part_row_begin[mid] <= r) lo = mid; else hi = mid - 1;
    }
    int p = lo;
    uint64_t dst = (uint64_t)part_win_begin[p] * 32ull + (cptr[r] - cptr[part_row_begin[p]]);
    uint64_t nchunks = cptr[r + 1] - cptr[r];
    int64_t a = load_index(crow, crow_dtype, r), e = load_index(crow, crow_dtype, r + 1);
    if (e < a) e = a;
    for (int64_t j = a + lane; j < e; j += 32) {
        int64_t c = load_index(col, col_dtype, j);
        if (c < 0 || c >= n_cols) { atomicExch(err, 2); continue; }  // leaves the sentinel in place
        uint64_t o = dst * 8ull + (uint64_t)(j - a);
        cols16[o] = (uint16_t)c;
        if constexpr (sizeof(VT) == 4) {
            vals[o] = load_value(val, val_dtype, j);
        } else if constexpr (sizeof(VT) == 2) {
            float v = load_value(val, val_dtype, j);
            if constexpr (std::is_same<VT, __half>::value) vals[o] = __float2half_rn(v);
            else vals[o] = __float2bfloat16_rn(v);
        }
    }
    if (lane == 0) {
        uint64_t t = dst + nchunks - 1;
        atomicOr(&tails[t >> 5], 1u << (uint32_t)(t & 31));
        row_chunk[r] = (uint32_t)dst;
        if (r == n_rows - 1) row_chunk[n_rows] = (uint32_t)(dst + nchunks);
    }
}

struct NoVal { char c; };

int build_ws_index(vs_index *idx, const void *d_crow, int crow_dtype, const void *d_col, int col_dtype,
                   const void *d_val, int val_dtype, cudaStream_t st) {
    const int64_t N = idx->n_rows;
    cudaDeviceProp prop;
    VS_CUDA(cudaGetDeviceProperties(&prop, idx->device));
    idx->n_ctas = prop.multiProcessorCount;
    idx->warps_per_cta = 32;
    idx->n_parts = idx->n_ctas * idx->warps_per_cta;
    const int P = idx->n_parts;

    uint64_t *d_cptr = nullptr, *d_nwin = nullptr;
    int *d_err = nullptr;
    void *d_tmp = nullptr;
    size_t tmp_bytes = 0;
    auto cleanup = [&]() { cudaFree(d_cptr); cudaFree(d_nwin); cudaFree(d_err); cudaFree(d_tmp); };

    VS_CUDA(cudaMalloc(&d_cptr, sizeof(uint64_t) * (size_t)(N + 1)));
    VS_CUDA(cudaMalloc(&d_nwin, sizeof(uint64_t)));
    VS_CUDA(cudaMalloc(&d_err, sizeof(int)));
    VS_CUDA(cudaMemsetAsync(d_err, 0, sizeof(int), st));
    {
        int64_t n = N + 1;
        row_chunks_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_crow, crow_dtype, N, d_cptr, d_err);
    }
    cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, d_cptr, d_cptr, (int64_t)(N + 1), st);
    VS_CUDA(cudaMalloc(&d_tmp, tmp_bytes ? tmp_bytes : 16));
    cub::DeviceScan::ExclusiveSum(d_tmp, tmp_bytes, d_cptr, d_cptr, (int64_t)(N + 1), st);

    VS_CUDA(cudaMalloc(&idx->part_row_begin, sizeof(uint32_t) * (size_t)(P + 1)));
    VS_CUDA(cudaMalloc(&idx->part_win_begin, sizeof(uint32_t) * (size_t)(P + 1)));
    part_rows_kernel<<<(P + 1 + 255) / 256, 256, 0, st>>>(d_cptr, N, P, idx->part_row_begin);
    part_windows_kernel<<<1, 32, 0, st>>>(d_cptr, idx->part_row_begin, P, idx->part_win_begin, d_nwin);
    uint64_t n_windows = 0;
    VS_CUDA(cudaMemcpyAsync(&n_windows, d_nwin, sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
    VS_CUDA(cudaStreamSynchronize(st));
    if (n_windows >= (1ull << 32)) { cleanup(); VS_REQUIRE(false, VS_ERR_UNSUPPORTED, "index too large for one shard: %llu windows", (unsigned long long)n_windows); }
    idx->n_windows = n_windows;

    const uint64_t n_chunks = n_windows * 32ull;
    size_t val_elem = 0;
    if (idx->kind == 1) val_elem = (idx->store_dtype == VS_F32) ? 4 : 2;
    VS_CUDA(cudaMalloc(&idx->cols, n_chunks ? n_chunks * 16 : 16));
    VS_CUDA(cudaMalloc(&idx->tails, n_windows ? n_windows * 4 : 4));
    VS_CUDA(cudaMalloc(&idx->row_chunk, sizeof(uint32_t) * (size_t)(N + 1)));
    if (val_elem) {
        VS_CUDA(cudaMalloc(&idx->vals, n_chunks ? n_chunks * 8 * val_elem : 16));
        VS_CUDA(cudaMemsetAsync(idx->vals, 0, n_chunks * 8 * val_elem, st));
    }
    VS_CUDA(cudaMemsetAsync(idx->tails, 0, n_windows * 4, st));
    VS_CUDA(cudaMemsetAsync(idx->row_chunk, 0, sizeof(uint32_t) * (size_t)(N + 1), st));
    const uint32_t sent = (uint32_t)idx->n_cols | ((uint32_t)idx->n_cols << 16);
    if (n_chunks) fill_u32_kernel<<<2048, 256, 0, st>>>((uint32_t *)idx->cols, n_chunks * 4, sent);

    if (N > 0) {
        unsigned blocks = (unsigned)((N * 32 + 255) / 256);
#define VS_FILL(VT)                                                                                              \
    fill_rows_kernel<VT><<<blocks, 256, 0, st>>>(d_crow, crow_dtype, d_col, col_dtype, d_val, val_dtype, N,      \
                                                 idx->n_cols, d_cptr, idx->part_row_begin, idx->part_win_begin,  \
                                                 P, (uint16_t *)idx->cols, (VT *)idx->vals, idx->tails,          \
                                                 idx->row_chunk, d_err)
        if (idx->kind == 2) VS_FILL(NoVal);
        else if (idx->store_dtype == VS_F32) VS_FILL(float);
        else if (idx->store_dtype == VS_F16) VS_FILL(__half);
        else VS_FILL(__nv_bfloat16);
#undef VS_FILL
    }
    int h_err = 0;
    VS_CUDA(cudaMemcpyAsync(&h_err, d_err, sizeof(int), cudaMemcpyDeviceToHost, st));
    VS_CUDA(cudaStreamSynchronize(st));
    VS_CUDA(cudaGetLastError());
    cleanup();
    VS_REQUIRE(h_err != 1, VS_ERR_INVALID, "crow_indices are not non-decreasing");
    VS_REQUIRE(h_err != 2, VS_ERR_INVALID, "col_indices outside [0, n_cols)");

    idx->stream_bytes = (int64_t)(n_chunks * (16 + 8 * val_elem) + n_windows * 4);
    idx->device_bytes = idx->stream_bytes + (int64_t)(sizeof(uint32_t) * (size_t)(N + 1 + 2 * (P + 1)));
    return VS_OK;
}

// ---- export: WS -> CSR (int64 crow/col, fp32 val), for SparseIndex.save (upstream index.py:181-202)
__global__ void export_len_kernel(const WsView idx, const uint32_t *row_chunk, uint64_t *len) {
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r > idx.n_rows) return;
    if (r == idx.n_rows) { len[r] = 0; return; }
    const uint16_t *c16 = (const uint16_t *)idx.cols;
    uint64_t a = (uint64_t)row_chunk[r] * 8ull;
    // the row's chunk count comes from its tail bit: walk chunks until the tail
    uint64_t ch = row_chunk[r];
    while (!((idx.tails[ch >> 5] >> (ch & 31)) & 1u)) ++ch;
    uint64_t e = (ch + 1) * 8ull;
    uint64_t n = 0;
    for (uint64_t j = a; j < e; ++j) n += (c16[j] != (uint16_t)idx.n_cols);
    len[r] = n;
}

__global__ void export_fill_kernel(const WsView idx, const uint32_t *row_chunk, const uint64_t *crow,
                                   int64_t *out_crow, int64_t *out_col, float *out_val) {
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r > idx.n_rows) return;
    out_crow[r] = (int64_t)crow[r];
    if (r == idx.n_rows) return;
    const uint16_t *c16 = (const uint16_t *)idx.cols;
    uint64_t a = (uint64_t)row_chunk[r] * 8ull;
    uint64_t o = crow[r], n = crow[r + 1] - crow[r];
    for (uint64_t j = 0, w = 0; w < n; ++j) {
        uint16_t c = c16[a + j];
        if (c == (uint16_t)idx.n_cols) continue;
        out_col[o + w] = c;
        float v = 1.0f;
        if (idx.kind == 1) {
            if (idx.store_dtype == VS_F32) v = ((const float *)idx.vals)[a + j];
            else if (idx.store_dtype == VS_F16) v = __half2float(((const __half *)idx.vals)[a + j]);
            else v = __bfloat162float(((const __nv_bfloat16 *)idx.vals)[a + j]);
        }
        out_val[o + w] = v;
        ++w;
    }
}

int export_ws_csr(const vs_index *idx, int64_t *d_crow, int64_t *d_col, float *d_val, cudaStream_t st) {
    const int64_t N = idx->n_rows;
    uint64_t *d_len = nullptr;
    void *d_tmp = nullptr;
    size_t tmp_bytes = 0;
    VS_CUDA(cudaMalloc(&d_len, sizeof(uint64_t) * (size_t)(N + 1)));
    unsigned blocks = (unsigned)((N + 1 + 255) / 256);
    export_len_kernel<<<blocks, 256, 0, st>>>(ws_view(idx), idx->row_chunk, d_len);
    cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, d_len, d_len, (int64_t)(N + 1), st);
    cudaError_t e = cudaMalloc(&d_tmp, tmp_bytes ? tmp_bytes : 16);
    if (e != cudaSuccess) { cudaFree(d_len); VS_CUDA(e); }
    cub::DeviceScan::ExclusiveSum(d_tmp, tmp_bytes, d_len, d_len, (int64_t)(N + 1), st);
    export_fill_kernel<<<blocks, 256, 0, st>>>(ws_view(idx), idx->row_chunk, d_len, d_crow, d_col, d_val);
    e = cudaStreamSynchronize(st);
    cudaFree(d_len);
    cudaFree(d_tmp);
    VS_CUDA(e);
    VS_CUDA(cudaGetLastError());
    return VS_OK;
}

}  // namespace vs
