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

}  // namespace vs

using namespace vs;

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
