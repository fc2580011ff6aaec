// sparsify.cu -- query sparsifier: keep the k largest activations of every row of a dense [B, V]
// fp32 embedding batch (plus, optionally, the columns of the row's own tokens), zero the rest, in place.
// Replaces upstream utils/sparse.py:8-19 (`build_topk_mask` = topk -> scatter a bool mask -> multiply) and the
// `logical_or(bow_mask, topk_mask)` of encoder/vdr.py:159-169.  One CTA per row: MSD radix select of the k-th largest
// rank key (ordered score << 32 | ~column: ties -> lower column, upstream's topk leaves them arbitrary) straight from
// the row in global memory, then one masking pass.  The sparse row feeds the inverted-list search, whose extract
// kernel compacts the survivors into (token, weight) lists.
#include "index.cuh"

namespace vs {

constexpr int kSparsifyThreads = 512;
constexpr int kMaxBow = 512;

__global__ void __launch_bounds__(kSparsifyThreads) sparsify_topk_kernel(float *q, int64_t ld, int n_cols, int k, const int32_t *bow_ids,
                                                                         int bow_ld, int bow_shift) {
    __shared__ uint32_t hist[256];
    __shared__ float s_bow_val[kMaxBow];
    __shared__ int s_bow_col[kMaxBow];
    const int tid = threadIdx.x, lane = tid & 31;
    float *row = q + (int64_t)blockIdx.x * ld;
    // the row's own tokens survive whatever their rank: remember their values
    int n_bow = 0;
    if (bow_ids != nullptr) {
        n_bow = bow_ld < kMaxBow ? bow_ld : kMaxBow;
        for (int j = tid; j < n_bow; j += kSparsifyThreads) {
            const int c = bow_ids[(int64_t)blockIdx.x * bow_ld + j] - bow_shift;
            s_bow_col[j] = (c >= 0 && c < n_cols) ? c : -1;
            s_bow_val[j] = (c >= 0 && c < n_cols) ? row[c] : 0.f;
        }
    }
    uint64_t kth = ~0ull;   // k == 0: nothing survives the rank test
    if (k >= n_cols) kth = 0ull;
    else if (k > 0) {
        uint64_t prefix = 0, mask = 0;
        int rem = k;
        for (int shift = 56; shift >= 0; shift -= 8) {
            for (int i = tid; i < 256; i += kSparsifyThreads) hist[i] = 0;
            __syncthreads();
            for (int base = 0; base < n_cols; base += kSparsifyThreads) {
                const int i = base + tid;
                const uint64_t x = (i < n_cols) ? make_key(row[i], (uint32_t)i) : 0ull;
                hist_add_aggregated(hist, (uint32_t)(x >> shift) & 255u, (i < n_cols) && ((x & mask) == prefix));
            }
            __syncthreads();
            uint32_t h[8], s = 0;
#pragma unroll
            for (int j = 0; j < 8; ++j) { h[j] = hist[lane * 8 + j]; s += h[j]; }
            uint32_t incl = s;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t o = __shfl_down_sync(0xffffffffu, incl, d);
                if (lane + d < 32) incl += o;
            }
            const uint32_t above = incl - s;
            const bool mine = (above < (uint32_t)rem) && ((uint32_t)rem <= incl);
            uint32_t digit = 0, newrem = 0;
            if (mine) {
                uint32_t acc = above;
#pragma unroll
                for (int j = 7; j >= 0; --j) {
                    if (acc < (uint32_t)rem && acc + h[j] >= (uint32_t)rem) { digit = lane * 8 + j; newrem = rem - acc; }
                    acc += h[j];
                }
            }
            const uint32_t owner = __ballot_sync(0xffffffffu, mine);
            const int src = __ffs(owner) - 1;
            digit = __shfl_sync(0xffffffffu, digit, src);
            newrem = __shfl_sync(0xffffffffu, newrem, src);
            prefix |= (uint64_t)digit << shift;
            mask |= (uint64_t)0xff << shift;
            rem = (int)newrem;
            __syncthreads();
        }
        kth = prefix;
    }
    __syncthreads();
    for (int i = tid; i < n_cols; i += kSparsifyThreads)
        if (make_key(row[i], (uint32_t)i) < kth) row[i] = 0.f;
    __syncthreads();
    for (int j = tid; j < n_bow; j += kSparsifyThreads)
        if (s_bow_col[j] >= 0) row[s_bow_col[j]] = s_bow_val[j];
}

// ---- dense [n_rows, ld] -> CSR: one warp per row, 32 columns per step (coalesced), the
// survivors of a step are ranked with one ballot.  Replaces `vectors.to_sparse_csr()` of the reference's build_index
// (retriever.py:299-305) and turns a sparsified query batch into the (token, weight) lists vs_search_sparse takes.
// Two calls: count (col == nullptr: row lengths into out[r]) and fill (crow given).
template <typename T>
__device__ __forceinline__ float dense_load(const T *p) { return (float)*p; }
template <>
__device__ __forceinline__ float dense_load<__half>(const __half *p) { return __half2float(*p); }
template <>
__device__ __forceinline__ float dense_load<__nv_bfloat16>(const __nv_bfloat16 *p) { return __bfloat162float(*p); }

template <typename T>
__global__ void __launch_bounds__(256) dense_to_csr_kernel(const T *x, int64_t n_rows, int64_t ld, int n_cols, int64_t *row_nnz_or_crow,
                                                           int32_t *col, float *val) {
    const int lane = threadIdx.x & 31;
    const int64_t r = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (r >= n_rows) return;
    const T *row = x + r * ld;
    const bool fill = col != nullptr;
    int64_t at = fill ? row_nnz_or_crow[r] : 0;
    int64_t n = 0;
    for (int c0 = 0; c0 < n_cols; c0 += 32) {
        const int c = c0 + lane;
        const float v = c < n_cols ? dense_load<T>(row + c) : 0.f;
        const bool nz = v != 0.0f;
        const uint32_t m = __ballot_sync(0xffffffffu, nz);
        if (fill && nz) {
            const int64_t o = at + n + __popc(m & ((1u << lane) - 1u));
            col[o] = c;
            val[o] = v;
        }
        n += __popc(m);
    }
    if (!fill && lane == 0) row_nnz_or_crow[r] = n;
}

}  // namespace vs

using namespace vs;

extern "C" int vs_dense_to_csr(int device, const void *d_x, int x_dtype, int64_t n_rows, int64_t ld, int n_cols,
                               int64_t *d_row_nnz_or_crow, int32_t *d_col, float *d_val, void *stream) {
    VS_REQUIRE(d_x != nullptr && n_rows >= 0 && n_cols > 0 && ld >= n_cols && d_row_nnz_or_crow != nullptr, VS_ERR_INVALID,
               "vs_dense_to_csr: bad argument");
    VS_REQUIRE(x_dtype == VS_F32 || x_dtype == VS_F16 || x_dtype == VS_BF16, VS_ERR_INVALID, "vs_dense_to_csr: f32 / f16 / bf16 input");
    VS_REQUIRE((d_col == nullptr) == (d_val == nullptr), VS_ERR_INVALID, "vs_dense_to_csr: d_col and d_val go together");
    if (n_rows == 0) return VS_OK;
    VS_CUDA(cudaSetDevice(device));
    const unsigned blocks = (unsigned)((n_rows + 7) / 8);
    cudaStream_t st = (cudaStream_t)stream;
    if (x_dtype == VS_F32) dense_to_csr_kernel<float><<<blocks, 256, 0, st>>>((const float *)d_x, n_rows, ld, n_cols, d_row_nnz_or_crow, d_col, d_val);
    else if (x_dtype == VS_F16) dense_to_csr_kernel<__half><<<blocks, 256, 0, st>>>((const __half *)d_x, n_rows, ld, n_cols, d_row_nnz_or_crow, d_col, d_val);
    else dense_to_csr_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>((const __nv_bfloat16 *)d_x, n_rows, ld, n_cols, d_row_nnz_or_crow, d_col, d_val);
    VS_CUDA(cudaGetLastError());
    return VS_OK;
}

extern "C" int vs_sparsify_topk(int device, float *d_q, int64_t B, int64_t ld, int n_cols, int k, const int32_t *d_bow_ids, int bow_ld,
                                int bow_shift, void *stream) {
    VS_REQUIRE(d_q != nullptr && B >= 0 && n_cols > 0 && ld >= n_cols && k >= 0, VS_ERR_INVALID, "vs_sparsify_topk: bad argument");
    VS_REQUIRE(d_bow_ids == nullptr || (bow_ld > 0 && bow_ld <= kMaxBow), VS_ERR_INVALID, "at most %d token ids per row", kMaxBow);
    if (B == 0) return VS_OK;
    VS_CUDA(cudaSetDevice(device));
    sparsify_topk_kernel<<<(unsigned)B, kSparsifyThreads, 0, (cudaStream_t)stream>>>(d_q, ld, n_cols, k, d_bow_ids, bow_ld, bow_shift);
    VS_CUDA(cudaGetLastError());
    return VS_OK;
}
